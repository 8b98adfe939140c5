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
        for m in re.finditer(r"^\s*([A-Za-z]+\.[A-Za-z0-9_.\[\]]+)\s*[=:]\s*(\S+)\s*$", open(tv).read(), flags=re.M):
            vectors[m.group(1)] = m.group(2)
    return params, d, vectors


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
    rho = P.with_auxsid(d["auxsid"].decode()).prefix().hex()
    if "der.rho" in vectors:
        assert vectors["der.rho"].lower() == rho


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
    assert rep["accepted"] and rep["shuffles"] == orep["shuffles"]


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
    (root / "vmnv_testvectors.txt").write_text("der.rho = %s\n" % P.with_auxsid("default").prefix().hex())
    monkeypatch.setattr(importlib.import_module(__name__), "REF", str(root))
    params2, d2, vectors = _load()
    assert marshalled_group(params2["pgroup"]) == (p, q, g) and d2 == d and "der.rho" in vectors
    rep = opr.verify_mix(G, _oracle_params(params2), 3, 2, d2)
    assert rep["accepted"]
