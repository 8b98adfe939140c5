"""Pin against a REAL Verificatum run, when an operator has supplied one (tests/golden/README.md): a proof
directory written by vmn (Java/GMP) must be accepted by the oracle's vmnv restatement and by the engine's, and
the values both derive on the way (global prefix, generators, seeds, challenges) must equal the test vectors
`vmnv -t` printed (mixnet/MixNetElGamalVerifyFiatShamirTool.java:82-224).  Skipped while tests/golden/nizkp_ref/
is absent -- until then parity with the Java/GMP path itself is unpinned."""
import importlib
import json
import os
import re

import pytest

from oracle import arithm as oar
from oracle import bytetree as bt
from oracle import protocols as opr
from oracle import testvectors

REF = os.path.join(os.path.dirname(__file__), "golden", "nizkp_ref")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nizkp")),
                               reason="no reference-produced proof directory (tests/golden/README.md)")


def _load():
    params = json.load(open(os.path.join(REF, "params.json")))
    d = {}
    root = os.path.join(REF, "nizkp")
    for base, _, files in os.walk(root):
        for fn in files:
            path = os.path.join(base, fn)
            d[os.path.relpath(path, root).replace(os.sep, "/")] = open(path, "rb").read()
    vectors = {}
    tv = os.path.join(REF, "vmnv_testvectors.txt")
    if os.path.exists(tv):
        vectors = testvectors.parse(open(tv).read())   # {(name, party or None): [values]}, the format of `vmnv -t`
    return params, d, vectors


def _check_vectors(ours, theirs):
    """Every scalar test vector we derive that the reference's dump also holds must agree (same party, same order
    of appearance).  Returns how many were compared."""
    seen, compared = {}, 0
    for name, party, value in ours:
        if name not in testvectors.SCALAR_NAMES:
            continue
        idx = seen.get((name, party), 0)
        seen[(name, party)] = idx + 1
        ref = theirs.get((name, party))
        if ref is None or idx >= len(ref):
            continue
        assert testvectors.same_value(name, value, ref[idx]), (name, party, value, ref[idx])
        compared += 1
    return compared


def marshalled_group(pgroup: str):
    """"<human readable>::<hex of node(leaf(class name), node(p, q, g, encoding))>" -> (p, q, g)."""
    t = bt.from_bytes(bytes.fromhex(pgroup.split("::")[-1]))
    assert t.children[0].value == b"com.verificatum.arithm.ModPGroup"
    p, q, g = (bt.bytes_to_int(x.value) for x in t.children[1].children[:3])
    return p, q, g


def _oracle_params(params):
    names = {"SHA-256": "sha256", "SHA-384": "sha384", "SHA-512": "sha512"}
    return opr.Params(vbitlenro=params["vbitlenro"], ebitlenro=params["ebitlenro"], rbitlen=params["rbitlen"],
                      rohash=names[params["rohash"]], prghash=names[params["prg"]], version=params["version"],
                      sid=params["sid"], pgroup_string=params["pgroup"])


@needs_ref
def test_oracle_accepts_the_reference_proof_directory():
    params, d, vectors = _load()
    G = oar.ModPGroup(*marshalled_group(params["pgroup"]))
    P = _oracle_params(params)
    rep = opr.verify_mix(G, P, params["k"], params["threshold"], d)
    assert rep["accepted"], rep
    # the global prefix, every seed and every challenge derived on the way against the `vmnv -t` dump
    if vectors:
        assert ("der.rho", None) in vectors, "the dump holds no der.rho: was vmnv run with -t par,der,bas,PoS,Dec,PoSC,CCPoS?"
        assert _check_vectors(rep["vectors"], vectors) >= 3


@needs_ref
@pytest.mark.gpu
def test_engine_accepts_the_reference_proof_directory(engine_cuda):
    params, d, _ = _load()
    vmx = engine_cuda
    vm = importlib.import_module("verificatum-vmn_b200.vmnv")
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G = vmx.arithm.ModPGroup(*marshalled_group(params["pgroup"]))
    sp = mix.SessionParams(vbitlenro=params["vbitlenro"], ebitlenro=params["ebitlenro"], rbitlen=params["rbitlen"],
                           rohash=params["rohash"], prghash=params["prg"], version=params["version"],
                           sid=params["sid"], pGroupString=params["pgroup"])
    V = vm.MixNetElGamalVerifyFiatShamirSession(G, sp, params["k"], params["threshold"])
    rep = V.verify(vm.ProofDirectory(d))
    orep = opr.verify_mix(oar.ModPGroup(*marshalled_group(params["pgroup"])), _oracle_params(params), params["k"],
                          params["threshold"], d)
    assert rep["accepted"] and rep["shuffles"] == orep["shuffles"] and rep["vectors"] == orep["vectors"]


def test_the_loader_reads_what_the_engine_writes(tmp_path, monkeypatch):
    """The consumer above on a directory of the same shape produced here (oracle mix, marshalled group string):
    proves the loader, the parameter mapping and the group parser before a real directory arrives."""
    from tests.cases import group_params
    from oracle.crypto import SeededRandomSource
    p, q, g = group_params(512)
    G = oar.ModPGroup(p, q, g)
    el = (p.bit_length() + 8) // 8
    hexgroup = bt.node(bt.leaf(b"com.verificatum.arithm.ModPGroup"),
                       bt.node(bt.int_leaf(p, el), bt.int_leaf(q, el), bt.int_leaf(g, el), bt.int32_leaf(1))).to_bytes().hex()
    params = {"version": "3.1.0", "sid": "SessionID", "k": 3, "threshold": 2, "vbitlenro": 256, "ebitlenro": 256,
              "rbitlen": 100, "prg": "SHA-256", "rohash": "SHA-256", "pgroup": "ModPGroup(test)::" + hexgroup}
    P = _oracle_params(params)
    rs = SeededRandomSource(b"\x07" * 32)
    x = oar.ring_random_element(G, SeededRandomSource(b"\x07" * 32), 100)
    pk = (G.g, G.op_exp(G.g, x))
    w = opr.demo_ciphertexts(G, pk, 7, SeededRandomSource(b"\x08" * 32))
    d, _ = opr.run_mix(G, P, 3, 2, w, rs, auxsid="default")
    root = tmp_path / "nizkp_ref"
    for name, data in d.items():
        f = root / "nizkp" / name
        f.parent.mkdir(parents=True, exist_ok=True)
        f.write_bytes(data)
    (root / "params.json").write_text(json.dumps(params))
    # the dump `vmnv -v -t ...` would print for it, in the reference's format, between other output
    honest = opr.verify_mix(G, P, 3, 2, d)
    dump = "Prepare to verify proof.\n" + testvectors.render(honest["vectors"]) + "\nVerification completed SUCCESSFULLY\n"
    (root / "vmnv_testvectors.txt").write_text(dump)
    monkeypatch.setattr(importlib.import_module(__name__), "REF", str(root))
    params2, d2, vectors = _load()
    assert marshalled_group(params2["pgroup"]) == (p, q, g) and d2 == d and ("der.rho", None) in vectors
    assert vectors[("PoS.s", 2)] and vectors[("Dec.v", None)] and len(vectors[("par.lambda", None)]) == 2
    rep = opr.verify_mix(G, _oracle_params(params2), 3, 2, d2)
    assert rep["accepted"]
    scalars = [v for v in rep["vectors"] if v[0] in testvectors.SCALAR_NAMES]
    assert _check_vectors(rep["vectors"], vectors) == len(scalars) >= 15
    # a dump that disagrees anywhere is caught (a challenge printed in hexadecimal is accepted)
    name, party, value = next(v for v in rep["vectors"] if v[0] == "PoS.v")
    vectors[(name, party)][0] = "%x" % int(value)
    assert _check_vectors(rep["vectors"], vectors) == len(scalars)
    vectors[(name, party)][0] = str(int(value) + 1)
    with pytest.raises(AssertionError):
        _check_vectors(rep["vectors"], vectors)
