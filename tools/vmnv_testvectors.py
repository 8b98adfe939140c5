"""Print the scalar test vectors of a proof directory in the format of the reference's `vmnv -t`
(oracle/testvectors.py), for a side-by-side diff with the output of a real Verificatum installation:

    vmnv -v -t par,der,bas,PoS,Dec,PoSC,CCPoS protInfo.xml nizkp > vmnv_testvectors.txt      # theirs
    python tools/vmnv_testvectors.py nizkp params.json > ours.txt                             # ours (CPU, the oracle)
    python tools/vmnv_testvectors.py nizkp params.json --compare vmnv_testvectors.txt         # or compared in place

params.json: the values of the protocol info file that enter the global prefix (tests/golden/README.md).
`--engine` runs the engine's verifier (the Python mirror over the C ABI: needs the B200, or VMX_LIBRARY_PATH set to the
emulation build) instead of the oracle; both record the same vectors.  Test infrastructure."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("nizkp")
    ap.add_argument("params")
    ap.add_argument("--engine", action="store_true")
    ap.add_argument("--compare", default="", help="a dump of `vmnv -t` to compare with instead of printing")
    args = ap.parse_args()
    from oracle import arithm as oar, protocols as opr, testvectors
    from tests.test_reference_pin import marshalled_group, _oracle_params, _check_vectors
    params = json.load(open(args.params))
    d = {}
    for base, _, files in os.walk(args.nizkp):
        for fn in files:
            path = os.path.join(base, fn)
            d[os.path.relpath(path, args.nizkp).replace(os.sep, "/")] = open(path, "rb").read()
    p, q, g = marshalled_group(params["pgroup"])
    if args.engine:
        vmx = importlib.import_module("verificatum-vmn_b200")
        vm = importlib.import_module("verificatum-vmn_b200.vmnv")
        mix = importlib.import_module("verificatum-vmn_b200.mixnet")
        sp = mix.SessionParams(vbitlenro=params["vbitlenro"], ebitlenro=params["ebitlenro"], rbitlen=params["rbitlen"],
                               rohash=params["rohash"], prghash=params["prg"], version=params["version"],
                               sid=params["sid"], pGroupString=params["pgroup"])
        V = vm.MixNetElGamalVerifyFiatShamirSession(vmx.arithm.ModPGroup(p, q, g), sp, params["k"], params["threshold"])
        try:
            rep = V.verify(vm.ProofDirectory(d))
        except vm.VerificationError as e:
            rep = dict(V.report, accepted=False, error=str(e))
    else:
        try:
            rep = opr.verify_mix(oar.ModPGroup(p, q, g), _oracle_params(params), params["k"], params["threshold"], d)
        except opr.MixVerificationError as e:
            raise SystemExit("fail-stop: %s" % e)
    if args.compare:
        n = _check_vectors(rep["vectors"], testvectors.parse(open(args.compare).read()))
        print("%d test vectors agree; accepted: %s" % (n, rep.get("accepted")))
    else:
        sys.stdout.write(testvectors.render(rep["vectors"]))
        print("\naccepted: %s" % rep.get("accepted"))


if __name__ == "__main__":
    main()
