"""Parity test bodies shared by the CPU (host-emulation) and GPU (CUDA) test modules: the engine
behind the C ABI against the oracle on the same seeded inputs, bit for bit."""
import dataclasses
import importlib
import random

import numpy as np

from oracle import arithm as oar
from oracle import bytetree as obt
from oracle import protocols as opr
from oracle.crypto import SeededRandomSource, PRGHeuristic as OPRG
from tests.cases import (EngineCase, OracleCase, col_values, elem_value, engine_elem, engine_group, group_params,
                         oracle_group, seed)


def _arrays(vmx, bits, n, rnd):
    A = vmx.arithm
    p, q, g = group_params(bits)
    G = A.ModPGroup(p, q, g)
    R = G.getPRing()
    xs = [pow(g, rnd.randrange(q), p) for _ in range(n)]
    es = [rnd.randrange(q) for _ in range(n)]
    X = G.toElementArray([A.PGroupElement(G, x) for x in xs])
    E = R.toElementArray([A.PFieldElement(R, e) for e in es])
    return A, G, R, p, q, g, xs, es, X, E


def group_ops(vmx, bits, n):
    rnd = random.Random(bits * 1000 + n)
    A, G, R, p, q, g, xs, es, X, E = _arrays(vmx, bits, n, rnd)
    OG = oar.ModPGroup(p, q, g)
    vals = lambda arr: [e.value for e in arr.elements()]
    assert vals(X) == xs and vals(E) == es                                     # codec round trip
    assert vals(X.mul(X)) == oar.g_mul(OG, xs, xs)
    assert vals(G.getg().exp(E)) == oar.g_exp(OG, g, es)                       # fixed base
    assert vals(X.exp(E)) == oar.g_exp(OG, xs, es)                             # variable base, per element
    s = A.PFieldElement(R, rnd.randrange(1 << 256))
    assert vals(X.exp(s)) == oar.g_exp(OG, xs, s.value)                        # variable base, one exponent
    assert X.expProd(E).value == oar.g_exp_prod(OG, xs, es)                    # multi-exponentiation
    short = [rnd.randrange(1 << 256) for _ in range(n)]
    S = R.toElementArray([A.PFieldElement(R, e) for e in short])
    assert X.expProd(S).value == oar.g_exp_prod(OG, xs, short)
    assert X.prod().value == oar.g_prod(OG, xs)
    perm = list(range(n))
    rnd.shuffle(perm)
    assert vals(X.permute(A.Permutation(perm))) == oar.permute(xs, perm)
    assert vals(X.shiftPush(G.getg())) == oar.shift_push(xs, g)
    keep = [i % 3 != 1 for i in range(n)]
    assert vals(X.extract(keep)) == [x for x, k in zip(xs, keep) if k]
    assert vals(X.copyOfRange(1, n - 1)) == xs[1:n - 1]
    assert X.get(n - 1).value == xs[-1]
    assert X.equals(X.copyOfRange(0, n)) and not X.equals(X.shiftPush(G.getg()))
    assert vals(X.inv()) == oar.g_inv(OG, xs)
    # simultaneous exponentiation x_i^s * y_i^{e_i} (the verifier's B-chain check): long and short exponents
    Yr = X.permute(A.Permutation(perm))
    ys = oar.permute(xs, perm)
    assert vals(X.expMulExp(s, Yr, E)) == [a * b % p for a, b in zip(oar.g_exp(OG, xs, s.value), oar.g_exp(OG, ys, es))]
    k613 = [rnd.randrange(1 << 613) % q for _ in range(n)]
    K = R.toElementArray([A.PFieldElement(R, e) for e in k613])
    assert vals(X.expMulExp(s, Yr, K)) == [a * b % p for a, b in zip(oar.g_exp(OG, xs, s.value), oar.g_exp(OG, ys, k613))]
    zero = A.PFieldElement(R, 0)
    assert vals(X.expMulExp(zero, Yr, K)) == oar.g_exp(OG, ys, k613)
    Y = X.mul(X)
    cols = G.expProd([X, Y], [5, -3], 3)
    assert vals(cols) == [pow(x, 5, p) * pow(pow(x * x % p, 3, p), -1, p) % p for x in xs]
    # single elements go through the engine as well
    a = A.PGroupElement(G, xs[0])
    assert a.exp(s).value == pow(xs[0], s.value, p)
    for e in (q - 5, q - 1, q - (1 << 200)):                                   # short negative exponents
        assert a.exp(A.PFieldElement(R, e)).value == pow(xs[0], e, p)
    assert A.PGroupElement(G, p - 1).exp(A.PFieldElement(R, q - 5)).value == pow(p - 1, q - 5, p)   # not a member
    assert a.mul(A.PGroupElement(G, xs[1])).value == xs[0] * xs[1] % p
    assert a.inv().value == pow(xs[0], -1, p)
    assert a.expMul(s, A.PGroupElement(G, xs[1])).value == pow(xs[0], s.value, p) * xs[1] % p


def edge_cases(vmx, bits):
    A = vmx.arithm
    p, q, g = group_params(bits)
    G = A.ModPGroup(p, q, g)
    R = G.getPRing()
    one = [A.PGroupElement(G, 1)] * 3
    X = G.toElementArray(one + [A.PGroupElement(G, p - 1 if pow(p - 1, q, p) == 1 else g)])
    zeros = R.toElementArray([A.PFieldElement(R, 0)] * 4)
    top = R.toElementArray([A.PFieldElement(R, q - 1)] * 4)
    assert [e.value for e in X.exp(zeros).elements()] == [1, 1, 1, 1]          # x^0
    assert X.expProd(zeros).value == 1
    assert [e.value for e in G.getg().exp(zeros).elements()] == [1] * 4
    assert [e.value for e in G.getg().exp(top).elements()] == [pow(g, q - 1, p)] * 4
    empty = G.toElementArray([])
    assert empty.size() == 0 and empty.to_matrix().shape == (0, G.elem_bytes)
    assert empty.inv().size() == 0
    # batched inversion: every shape of the product tree (arity 4), units and p - 1 among the elements;
    # simultaneous exponentiation with a zero scalar / zero array and on the empty array
    rnd = random.Random(bits)
    for m in (1, 2, 3, 4, 5, 16, 17, 21, 65):
        vs = [pow(g, rnd.randrange(1, q), p) for _ in range(m)]
        vs[rnd.randrange(m)] = 1
        V = G.toElementArray([A.PGroupElement(G, v) for v in vs])
        assert [e.value for e in V.inv().elements()] == [pow(v, -1, p) for v in vs], m
    assert empty.expMulExp(A.PFieldElement(R, 3), empty, R.toElementArray([])).size() == 0
    assert [e.value for e in X.expMulExp(A.PFieldElement(R, 0), X, zeros).elements()] == [1, 1, 1, 1]
    assert [e.value for e in X.expMulExp(A.PFieldElement(R, 2), X, top).elements()] == \
        [pow(e.value, q + 1, p) for e in X.elements()]
    single = G.toElementArray([A.PGroupElement(G, g)])
    e1 = R.toElementArray([A.PFieldElement(R, 5)])
    assert single.expProd(e1).value == pow(g, 5, p) and single.prod().value == g
    # range and membership violations are ArithmFormatException, not crashes
    for bad in (0, p, p + 1):
        try:
            m = np.frombuffer(bad.to_bytes(G.elem_bytes, "big"), dtype=np.uint8)
            G.toElementArray(1, m)
            assert False, bad
        except A.ArithmFormatException:
            pass
    nonmember = next(v for v in range(2, 50) if pow(v, q, p) != 1)
    try:
        G.toElementArray(1, np.frombuffer(nonmember.to_bytes(G.elem_bytes, "big"), dtype=np.uint8))
        assert False
    except A.ArithmFormatException:
        pass
    try:
        R.toElementArray(1, vmx.eio.ByteTreeReader(vmx.eio.ByteTreeContainer(
            vmx.eio.ByteTreeLeaf(q.to_bytes(G.ring_bytes, "big"))).to_bytes()))
        assert False
    except A.ArithmFormatException:
        pass
    # membership on import = Legendre symbol (binary Jacobi kernel) for safe primes: arrays of residues
    # pass, one non-residue anywhere rejects, element by element against Euler's criterion
    import random as _r
    rq = _r.Random(1000 + bits)
    cand = [rq.randrange(1, p) for _ in range(40)] + [1, p - 1, 2, 3, 4, p - 2, (p - 1) // 2, (p + 1) // 2]
    qr = [v for v in cand if pow(v, q, p) == 1]
    nqr = [v for v in cand if pow(v, q, p) != 1]
    assert qr and nqr
    as_m = lambda vs: np.frombuffer(b"".join(v.to_bytes(G.elem_bytes, "big") for v in vs), dtype=np.uint8)
    ok_arr = G.toElementArray(len(qr), as_m(qr), check_membership=True)
    assert [e.value for e in ok_arr.elements()] == qr
    for v in nqr:
        for pos in (0, len(qr) // 2, len(qr)):
            try:
                G.toElementArray(len(qr) + 1, as_m(qr[:pos] + [v] + qr[pos:]), check_membership=True)
                assert False, "non-residue accepted"
            except A.ArithmFormatException:
                pass
    # single-element inversion (host binary Euclid in the engine) against Python, incl. 1 and p-1
    rnd = random.Random(bits)
    for v in [1, p - 1, 2, g, p - 2] + [rnd.randrange(1, p) for _ in range(12)]:
        assert A.PGroupElement(G, v).inv().value == pow(v, -1, p)
    assert [e.value for e in X.inv().elements()] == [pow(e.value, -1, p) for e in X.elements()]
    # size mismatch is an error status, not a crash
    try:
        X.mul(single)
        assert False
    except vmx._native.VmxError as e:
        assert e.status == vmx._native.VMX_ESIZE


def ring_ops(vmx, bits, n):
    rnd = random.Random((bits if isinstance(bits, int) else 256) * 7 + n)
    A = vmx.arithm
    G, OG = engine_group(vmx, bits), oracle_group(bits)
    q = OG.q
    R = G.getPRing()
    a = [rnd.randrange(q) for _ in range(n)]
    b = [rnd.randrange(q) for _ in range(n)]
    Ar = R.toElementArray([A.PFieldElement(R, v) for v in a])
    Br = R.toElementArray([A.PFieldElement(R, v) for v in b])
    vals = lambda arr: [e.value for e in arr.elements()]
    assert vals(Ar.add(Br)) == [(x + y) % q for x, y in zip(a, b)]
    assert vals(Ar.neg()) == [(-x) % q for x in a]
    assert vals(Ar.mul(Br)) == [x * y % q for x, y in zip(a, b)]
    sc = A.PFieldElement(R, rnd.randrange(q))
    assert vals(Ar.mulAdd(sc, Br)) == [(x * sc.value + y) % q for x, y in zip(a, b)]
    assert Ar.innerProduct(Br).value == oar.r_inner(OG, a, b)
    assert Ar.sum().value == sum(a) % q
    pr = 1
    for v in a:
        pr = pr * v % q
    assert Ar.prod().value == pr
    assert vals(Ar.prods()) == oar.r_prods(OG, a)
    x, d = Br.recLin(Ar)
    ox, od = oar.r_rec_lin(OG, b, a)
    assert vals(x) == ox and d.value == od
    perm = list(range(n))
    rnd.shuffle(perm)
    assert vals(Ar.permute(A.Permutation(perm))) == oar.permute(a, perm)
    assert vals(Ar.shiftPush(sc)) == [sc.value] + a[:-1]
    assert Ar.bitLength() == max(v.bit_length() for v in a)


def random_sources(vmx, bits, n):
    """Same seed => same arrays as the oracle's reading of the VCR sampling rules, at any stream offset."""
    A, cr = vmx.arithm, vmx.crypto
    p, q, g = group_params(bits)
    G = A.ModPGroup(p, q, g)
    R = G.getPRing()
    OG = oar.ModPGroup(p, q, g)
    rs = cr.PRGHeuristic()
    rs.setSeed(seed("rs"))
    ors = SeededRandomSource(seed("rs"))
    vals = lambda arr: [e.value for e in arr.elements()]
    assert vals(R.randomElementArray(n, rs, 100)) == oar.ring_random_array(OG, n, ors, 100)
    assert R.randomElement(rs, 100).value == oar.ring_random_element(OG, ors, 100)        # odd offset from here on
    assert vals(R.toElementArray(A.LargeIntegerArray.random(n, 612, rs, R))) == [v % q for v in oar.lia_random(n, 612, ors)]
    assert vals(G.randomElementArray(n, rs, 100)) == oar.group_random_array(OG, n, ors, 100)
    assert vals(R.randomElementArray(n, rs, 100)) == oar.ring_random_array(OG, n, ors, 100)
    assert list(A.Permutation.random(n, rs, 100).table) == oar.permutation_random(n, ors, 100)
    # long requests are expanded on the device (vmx_prg_bytes_sha256): same permutation, same stream position
    big = 700
    assert list(A.Permutation.random(big, rs, 100, G).table) == oar.permutation_random(big, ors, 100)
    assert rs.getBytes(50) == ors.get_bytes(50)
    # n >= 4096: the keys are also RANKED on the device (vmx_permutation_prg_sha256); short keys (statDist 0)
    # tie and are ordered by index, as a stable sort does
    for big, sd in ((5000, 100), (4100, 0)):
        assert list(A.Permutation.random(big, rs, sd, G).table) == oar.permutation_random(big, ors, sd)
    assert rs.getBytes(3) == ors.get_bytes(3)
    prg = cr.PRGHeuristic()
    prg.setSeed(seed("batch"))
    assert vals(R.toElementArray(A.LargeIntegerArray.random(n, 256, prg, R))) == opr.batch_vector("sha256", seed("batch"), n, 256)


def transcript_parity(vmx, bits, n):
    """The engine's mix-server and the oracle's, fed the same seeds, publish identical bytes; each
    verifier accepts the other's proof; a corrupted proof is rejected by both."""
    oc = OracleCase(bits, n)
    ec = EngineCase(vmx, bits, n)
    assert col_values(ec.w) == oc.w and ec.x.value == oc.x
    prover = ec.session("prover")
    h = prover.deriveGenerators(n)
    assert col_values(h) == oc.h
    proof, out = prover.shuffle(1, ec.w, generators=h, keep_output=True)
    owp, oproof = opr.shuffle_and_prove(oc.G, oc.params, oc.pk, oc.w, oc.h, SeededRandomSource(seed("prover")))
    assert col_values(out) == owp
    assert proof.output == oproof["output"]
    assert proof.permutationCommitment == oproof["permutationCommitment"]
    assert proof.commitment == oproof["commitment"]
    assert proof.reply == oproof["reply"]
    # cross verification
    verifier = ec.session(None)
    ok, out2 = verifier.verify(1, ec.w, proof, generators=h)
    assert ok and col_values(out2) == owp
    # online verification (the verifier hashes each message as the prover publishes it): same transcript, same
    # verdict; a message altered between the board and the final proof falls back to the offline order
    prover2 = ec.session("prover")
    ov = verifier.beginVerify(1, ec.w, generators=h)
    proof2, _ = prover2.shuffle(1, ec.w, generators=h, publish=ov.publish)
    assert dataclasses.astuple(proof2) == dataclasses.astuple(proof)
    ok2, out3 = ov.finish(proof2)
    assert ok2 and col_values(out3) == owp
    ov = verifier.beginVerify(1, ec.w, generators=h)
    seen = {}
    prover3 = ec.session("prover")
    proof3, _ = prover3.shuffle(1, ec.w, generators=h, publish=lambda nm, msg: (seen.__setitem__(nm, msg), ov.publish(nm, msg)))
    assert set(seen) == {"output", "permutationCommitment", "commitment", "reply"}
    raw = bytearray(proof3.commitment)
    raw[len(raw) // 2] ^= 0x04
    okb, outb = ov.finish(dataclasses.replace(proof3, commitment=bytes(raw)))
    assert okb is False and col_values(outb) == oc.w
    # the output on the board differs from the one the proof is finally presented with (same length): the streamed
    # seed hash covered other bytes and must not be used -- same verdicts as the offline order, both ways round
    other_out = bytearray(proof.output)
    other_out[len(other_out) // 3] ^= 0x01
    other_out = bytes(other_out)
    for board, final in ((other_out, proof), (proof.output, dataclasses.replace(proof, output=other_out))):
        ov = verifier.beginVerify(1, ec.w, generators=h)
        ov.publish("output", board)
        ov.publish("permutationCommitment", proof.permutationCommitment)
        ov.publish("commitment", proof.commitment)
        got = ov.finish(final)
        want = verifier.verify(1, ec.w, final, generators=h)
        assert got[0] == want[0] and col_values(got[1]) == col_values(want[1])
    assert verifier.verify(1, ec.w, proof, generators=h)[0] is True
    ov = verifier.beginVerify(1, ec.w, generators=h)
    ov.publish("output", proof.output)
    ov.publish("permutationCommitment", proof.permutationCommitment[:-2])      # a truncated message on the board
    ov.publish("commitment", proof.commitment)
    assert ov.finish(proof)[0] is True                                          # ... the proof itself is intact
    ov = verifier.beginVerify(1, ec.w, generators=h)
    ov.publish("output", proof.output)
    ov.publish("permutationCommitment", bytes(proof.permutationCommitment) + b"\x00")   # trailing byte: not canonical
    ov.publish("commitment", proof.commitment)
    trailing = dataclasses.replace(proof, permutationCommitment=bytes(proof.permutationCommitment) + b"\x00")
    assert ov.finish(trailing)[0] == verifier.verify(1, ec.w, trailing, generators=h)[0]
    assert opr.verify_shuffle(oc.G, oc.params, oc.pk, oc.w, oc.h, dataclasses.asdict(proof))
    # corrupted proofs: same verdict from both
    for field in ("reply", "commitment", "permutationCommitment", "output"):
        raw = bytearray(getattr(proof, field))
        raw[len(raw) // 2] ^= 0x04
        bad = dataclasses.replace(proof, **{field: bytes(raw)})
        okb, outb = verifier.verify(1, ec.w, bad, generators=h)
        assert okb is False and col_values(outb) == oc.w          # "Replacing output with input"
        assert opr.verify_shuffle(oc.G, oc.params, oc.pk, oc.w, oc.h, dataclasses.asdict(bad)) is False
    for field in ("reply", "commitment", "permutationCommitment", "output"):
        bad = dataclasses.replace(proof, **{field: getattr(proof, field)[:-2]})
        assert verifier.verify(1, ec.w, bad, generators=h)[0] is False


def accept_reject_properties(vmx, bits, n):
    """Size-independent properties at sizes the oracle cannot follow: an honest proof verifies,
    decrypting the output gives a permutation of the decrypted input (checked through products:
    prod(dec(w')) == prod(dec(w))), a single flipped limb is rejected."""
    ec = EngineCase(vmx, bits, n, label="big")
    prover = ec.session("prover-big")
    h = prover.deriveGenerators(n)
    proof, out = prover.shuffle(1, ec.w, generators=h, keep_output=True)
    verifier = ec.session(None)
    ok, out2 = verifier.verify(1, ec.w, proof, generators=h)
    assert ok and out2.equals(out)
    dec = lambda w: w.project(1).prod().mul(w.project(0).prod().exp(ec.x).inv())
    assert dec(out).equals(dec(ec.w))
    raw = bytearray(proof.reply)
    raw[-7] ^= 0x10
    assert verifier.verify(1, ec.w, dataclasses.replace(proof, reply=bytes(raw)), generators=h)[0] is False


def _bt_bytes(t) -> bytes:
    return t.to_bytes()


def _commitment_instance(vmx, bits, n, label):
    """u = (h * g^r) permuted, built on both sides from the same seeds (mixnet/PermutationCommitment.java:191-215)."""
    A = vmx.arithm
    cr = vmx.crypto
    oc = OracleCase(bits, n, label)
    ec = EngineCase(vmx, bits, n, label)
    OG = oc.G
    h = ec.session(None).deriveGenerators(n)
    rs = cr.PRGHeuristic()
    rs.setSeed(seed(label + "/commit"))
    ors = SeededRandomSource(seed(label + "/commit"))
    r = ec.G.getPRing().randomElementArray(n, rs, 100)
    pi = A.Permutation.random(n, rs, 100)
    o_r = oar.ring_random_array(OG, n, ors, 100)
    o_pi = oar.permutation_random(n, ors, 100)
    tmp = ec.G.getg().exp(r)
    tmp2 = h.mul(tmp)
    u = tmp2.permute(pi)
    o_u = oar.permute(oar.g_mul(OG, oc.h, oar.g_exp(OG, OG.g, o_r)), o_pi)
    assert col_values(u) == o_u
    return oc, ec, h, (r, pi, u), (o_r, o_pi, o_u), rs, ors


def posc_parity(vmx, bits, n):
    """PoSCBasicTW (hvzk/PoSCBasicTW.java): identical commitment and reply bytes, cross verification,
    rejection of a corrupted reply (what hvzk/TestPoSCBasicTW.java:147-163 asserts) on both sides."""
    hv = importlib.import_module("verificatum-vmn_b200.hvzk")
    cr = vmx.crypto
    oc, ec, h, (r, pi, u), (o_r, o_pi, o_u), rs, ors = _commitment_instance(vmx, bits, n, "posc")
    OG = oc.G
    P = hv.PoSCBasicTW(256, 256, 100, cr.PRGHeuristic(), rs)
    P.setInstance(ec.G.getg(), h, u, r, pi)
    OP = opr.PoSCBasicTW(OG, 256, 256, 100, "sha256", ors)
    OP.set_instance(OG.g, oc.h, o_u, o_r, o_pi)
    c = _bt_bytes(P.commit(seed("posc/batch")))
    oc_ = OP.commit(seed("posc/batch")).to_bytes()
    assert c == oc_
    v = int.from_bytes(seed("posc/challenge"), "big")
    rep = _bt_bytes(P.reply(v))
    assert rep == OP.reply(v).to_bytes()
    # engine verifier on the oracle's bytes and vice versa
    V = hv.PoSCBasicTW(256, 256, 100, cr.PRGHeuristic(), None)
    V.setInstance(ec.G.getg(), h, u)
    V.setBatchVector(seed("posc/batch"))
    assert _bt_bytes(V.setCommitment(vmx.eio.ByteTreeReader(oc_))) == oc_
    V.setChallenge(v)
    assert V.verify(vmx.eio.ByteTreeReader(rep)) is True
    OV = opr.PoSCBasicTW(OG, 256, 256, 100, "sha256", None)
    OV.set_instance(OG.g, oc.h, o_u)
    OV.set_batch_vector(seed("posc/batch"))
    OV.set_commitment(obt.from_bytes(c))
    OV.set_challenge(v)
    assert OV.verify(obt.from_bytes(rep)) is True
    bad = bytearray(rep)
    bad[len(bad) // 3] ^= 1
    V.setBatchVector(seed("posc/batch"))
    assert V.verify(vmx.eio.ByteTreeReader(bytes(bad))) is False
    assert OV.verify(obt.from_bytes(bytes(bad))) is False


def ccpos_parity(vmx, bits, n):
    """CCPoSBasicW (hvzk/CCPoSBasicW.java): commitment-consistent proof of a shuffle."""
    A = vmx.arithm
    hv = importlib.import_module("verificatum-vmn_b200.hvzk")
    cr = vmx.crypto
    oc, ec, h, (r, pi, u), (o_r, o_pi, o_u), rs, ors = _commitment_instance(vmx, bits, n, "ccpos")
    OG = oc.G
    # re-encrypt and permute with the committed permutation (mixnet/ShufflerElGamalSession.java:771-795)
    s = ec.G.getPRing().randomElementArray(n, rs, 100)
    o_s = oar.ring_random_array(OG, n, ors, 100)
    factors = ec.pk.exp(s)
    reenc = ec.w.mul(factors)
    piinv = pi.inv()
    wp = reenc.permute(piinv)
    o_wp = oar.permute(oar.g_mul(OG, oc.w, oar.g_exp(OG, oc.pk, o_s)), oar.perm_inv(o_pi))
    assert col_values(wp) == o_wp
    P = hv.CCPoSBasicW(256, 256, 100, cr.PRGHeuristic())
    P.setInstance(ec.G.getg(), h, u, ec.pk, ec.w, wp, r, pi, s)
    OP = opr.CCPoSBasicW(OG, 256, 256, 100, "sha256")
    OP.set_instance(OG.g, oc.h, o_u, oc.pk, oc.w, o_wp, o_r, o_pi, o_s)
    c = _bt_bytes(P.commit(seed("ccpos/batch"), rs))
    oc_ = OP.commit(seed("ccpos/batch"), ors).to_bytes()
    assert c == oc_
    v = int.from_bytes(seed("ccpos/challenge"), "big")
    rep = _bt_bytes(P.reply(v))
    assert rep == OP.reply(v).to_bytes()
    V = hv.CCPoSBasicW(256, 256, 100, cr.PRGHeuristic())
    V.setInstance(ec.G.getg(), h, u, ec.pk, ec.w, wp)
    V.setBatchVector(seed("ccpos/batch"))
    V.computeAB()
    V.setCommitment(vmx.eio.ByteTreeReader(oc_))
    V.setChallenge(v)
    assert V.verify(vmx.eio.ByteTreeReader(rep)) is True
    OV = opr.CCPoSBasicW(OG, 256, 256, 100, "sha256")
    OV.set_instance(OG.g, oc.h, o_u, oc.pk, oc.w, o_wp)
    OV.set_batch_vector(seed("ccpos/batch"))
    OV.compute_AB()
    OV.set_commitment(obt.from_bytes(c))
    OV.set_challenge(v)
    assert OV.verify(obt.from_bytes(rep)) is True
    bad = bytearray(rep)
    bad[-3] ^= 0x20
    assert V.verify(vmx.eio.ByteTreeReader(bytes(bad))) is False
    assert OV.verify(obt.from_bytes(bytes(bad))) is False


def decryption_parity(vmx, bits, n, k=3, threshold=2):
    """Decryption factors, their combination with modified Lagrange coefficients and the batched proof
    (elgamal/DistrElGamalSession.java:377-406, elgamal/DistrElGamalSessionBasic.java:465-727), k parties,
    threshold of them combined; the plaintexts g^{m} come back."""
    A = vmx.arithm
    eg = importlib.import_module("verificatum-vmn_b200.elgamal")
    cr = vmx.crypto
    oc = OracleCase(bits, n, "dec")
    ec = EngineCase(vmx, bits, n, "dec")
    OG, G = oc.G, ec.G
    q = OG.q
    R = G.getPRing()
    # Shamir shares of the secret key oc.x: polynomial of degree threshold-1, x_l = f(l)
    ors = SeededRandomSource(seed("dec/poly"))
    coeffs = [oc.x] + [oar.ring_random_element(OG, ors, 100) for _ in range(threshold - 1)]
    xs = {l: sum(c * pow(l, i, q) for i, c in enumerate(coeffs)) % q for l in range(1, k + 1)}
    ys = {l: OG.op_exp(OG.g, xs[l]) for l in range(1, k + 1)}
    correct = [False] + [True] * k
    ints = opr.modified_lagrange_coefficients(q, correct, k, threshold)
    assert ints == eg.modifiedLagrangeCoefficients(R, correct, k, threshold)
    u, o_u = ec.w.project(0), oc.w[0]
    # decryption factors of every party
    f, o_f = {}, {}
    inv_factor = pow(opr.prod_factor(q, k), -1, q)
    for l in range(1, k + 1):
        f[l] = eg.decryptionFactors(u, A.PFieldElement(R, xs[l]), k)
        o_f[l] = opr.decryption_factors(OG, o_u, xs[l], inv_factor)
        assert col_values(f[l]) == o_f[l]
    combined = eg.combineDecryptionFactors(f, correct, k, threshold)
    o_combined = opr.combine_decryption_factors(OG, o_f, correct, k, threshold)
    assert col_values(combined) == o_combined
    # plaintexts = v * combined  (mixnet/MixNetElGamalVerifyFiatShamirSession.java:1267-1275); the
    # modified coefficients carry the factor prodFactor, cancelled by inverseFactor in the exponent
    plain = ec.w.project(1).mul(combined)
    o_plain = [OG.op_mul(b, OG.op_inv(OG.op_exp(a, oc.x))) for a, b in zip(oc.w[0], oc.w[1])]
    assert col_values(plain) == o_plain
    # the proof of party j, verified by everybody (identical bytes on both sides)
    v = int.from_bytes(seed("dec/challenge"), "big")
    engines, oracles = {}, {}
    g_el = G.getg()
    y_el = {l: engine_elem(vmx, G, ys[l]) for l in ys}
    for j in range(1, k + 1):
        E = eg.DistrElGamalSessionBasic(j, k, threshold, 256, 100, cr.PRGHeuristic())
        E.setInstance(g_el, u, y_el, f, A.PFieldElement(R, xs[j]))
        E.setBatchVector(seed("dec/batch"))
        E.batchInput()
        O = opr.DistrElGamalSessionBasic(OG, j, k, threshold, 256, 100, "sha256", OG.g, ys, o_u, xs[j])
        O.f = o_f
        O.set_batch_vector(seed("dec/batch"))
        O.batch_input()
        assert elem_value(E.A) == O.A
        rs = cr.PRGHeuristic()
        rs.setSeed(seed("dec/commit%d" % j))
        c = E.commit(rs).to_bytes()
        assert c == O.commit(SeededRandomSource(seed("dec/commit%d" % j))).to_bytes()
        rep = E.reply(v).to_bytes()
        assert rep == O.reply(v).to_bytes()
        engines[j], oracles[j] = (E, c, rep), (O, c, rep)
    # party 1 verifies everybody individually and combined
    E1, O1 = engines[1][0], oracles[1][0]
    for l in range(2, k + 1):
        E1.setCommitment(l, vmx.eio.ByteTreeReader(engines[l][1]))
        E1.setReply(l, vmx.eio.ByteTreeReader(engines[l][2]))
        O1.set_commitment(l, obt.from_bytes(engines[l][1]))
        O1.set_reply(l, obt.from_bytes(engines[l][2]))
    for l in range(1, k + 1):
        E1.batch(l)
        O1.batch(l)
        assert elem_value(E1.B[l]) == O1.B[l]
        assert E1.verify(l, v) is True and O1.verify(l, v) is True
    E1.combine(correct)
    O1.combine(correct)
    # combinedy = the joint public key y = g^x (elgamal/DistrElGamalSession.java:422-428)
    E1.combinedy, E1.combinedf = engine_elem(vmx, G, oc.pk[1]), combined
    E1.batchCombined()
    O1.batch_combined(o_combined)
    assert elem_value(E1.combinedB) == O1.combinedB
    assert E1.verifyCombined(v) is True and O1.verify_combined(oc.pk[1], v) is True
    # a wrong reply is rejected by both
    E1.k_x[k] = E1.k_x[k].add(R.getONE())
    O1.k_x[k] = (O1.k_x[k] + 1) % q
    assert E1.verify(k, v) is False and O1.verify(k, v) is False


def committed_shuffle_parity(vmx, bits, maxciph, n):
    """Pre-computation with maxciph generators (permutation commitment + PoSC), shrink to n actual
    ciphertexts (keep list, Permutation.shrink, extract), commitment-consistent shuffle (CCPoS):
    the engine's and the oracle's published bytes are identical and each verifies the other's."""
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    oc = OracleCase(bits, n, "committed")
    ec = EngineCase(vmx, bits, n, "committed")
    OG = oc.G
    oh = opr.independent_generators(OG, "sha256", oc.params.prefix(), "generators", maxciph, oc.params.rbitlen)
    ostate, opub = opr.precomp(OG, oc.params, oc.pk, oh, SeededRandomSource(seed("committed/prover")))
    okeep = opr.shrink(OG, ostate, n)
    owp, oproof = opr.committed_shuffle(OG, oc.params, oc.pk, ostate, oc.w, SeededRandomSource(seed("committed/prove2")))
    # engine prover
    prover = ec.session("committed/prover")
    cs = mix.CommittedShuffler(prover, 1, maxciph)
    assert col_values(cs.generators) == oh
    pub = cs.precomp()
    assert pub == opub
    keep = cs.shrink(n)
    assert keep == okeep
    assert list(cs.permutationCommitment.permutation.table) == ostate["pi"]
    assert col_values(cs.permutationCommitment.commitment) == ostate["u"]
    prover.randomSource.setSeed(seed("committed/prove2"))
    proof, out = cs.shuffle(ec.w, keep_output=True)
    assert col_values(out) == owp
    assert dataclasses.asdict(proof) == oproof
    # engine verifier: PoSC on the full commitment, shrink with the published keep list, CCPoS
    verifier = ec.session(None)
    gens = verifier.deriveGenerators(maxciph)
    pcv = mix.PermutationCommitment(verifier, gens)
    assert pcv.verify(*opub) is True
    assert pcv.shrink(n, okeep) == okeep
    sgens = gens.copyOfRange(0, n)
    ok, out2 = mix.verifyCommittedShuffle(verifier, 1, sgens, pcv.commitment, ec.w, proof)
    assert ok is True and col_values(out2) == owp
    # online verification (the verifier's seed hash starts when the output is published, beside the prover's): same
    # verdict and output; a board message other than the output finally presented is hashed again, not trusted
    prover.randomSource.setSeed(seed("committed/prove2"))
    ov = mix.OnlineCommittedVerification(verifier, 1, sgens, pcv.commitment, ec.w)
    proof_on, _ = cs.shuffle(ec.w, publish=ov.publish)
    assert dataclasses.asdict(proof_on) == oproof
    ok_on, out_on = ov.finish(proof_on)
    assert ok_on is True and col_values(out_on) == owp
    ov = mix.OnlineCommittedVerification(verifier, 1, sgens, pcv.commitment, ec.w)
    other = bytearray(proof.output)
    other[len(other) // 2] ^= 1
    ov.publish("output", bytes(other))
    ok_on, out_on = ov.finish(proof)
    assert ok_on is True and col_values(out_on) == owp
    ov = mix.OnlineCommittedVerification(verifier, 1, sgens, pcv.commitment, ec.w)
    ov.publish("output", proof.output)
    ok_on, out_on = ov.finish(dataclasses.replace(proof, output=bytes(other)))
    want = mix.verifyCommittedShuffle(verifier, 1, sgens, pcv.commitment, ec.w, dataclasses.replace(proof, output=bytes(other)))
    assert ok_on is want[0] is False and col_values(out_on) == col_values(want[1]) == oc.w
    # oracle verifier on the engine's bytes
    ou = oar.parse_array(OG, obt.from_bytes(pub[0]), maxciph)
    assert opr.posc_verify(OG, oc.params, OG.g, oh, ou, pub[1], pub[2]) is True
    assert opr.ccpos_verify(OG, oc.params, OG.g, oh[:n], ostate["u"], oc.pk, oc.w, owp, proof.commitment, proof.reply)
    # corruption: a flipped bit in the CCPoS reply is rejected by both and the output replaced by the input
    raw = bytearray(proof.reply)
    raw[-2] ^= 8
    okb, outb = mix.verifyCommittedShuffle(verifier, 1, sgens, pcv.commitment, ec.w,
                                           dataclasses.replace(proof, reply=bytes(raw)))
    assert okb is False and col_values(outb) == oc.w
    assert opr.ccpos_verify(OG, oc.params, OG.g, oh[:n], ostate["u"], oc.pk, oc.w, owp, proof.commitment, bytes(raw)) is False
    # a corrupted PoSC reply makes the commitment trivial (the generators)
    rawp = bytearray(pub[2])
    rawp[7] ^= 1
    pcb = mix.PermutationCommitment(verifier, gens)
    assert pcb.verify(pub[0], pub[1], bytes(rawp)) is False and col_values(pcb.commitment) == oh
    assert opr.posc_verify(OG, oc.params, OG.g, oh, ou, pub[1], bytes(rawp)) is False
    # a keep list with the wrong number of entries is replaced by the trivial one
    bad_keep = bytearray(okeep)
    bad_keep[-1] ^= 1
    pcc = mix.PermutationCommitment(verifier, gens)
    pcc.verify(*opub)
    assert pcc.shrink(n, bytes(bad_keep)) == vmx.eio.booleanArrayToByteTree([i < n for i in range(maxciph)]).to_bytes()


def ec_group_ops(vmx, curve, n):
    """ECqPGroup arrays behind the C ABI against oracle/ec.py: every array method, the unit element and
    coincident / opposite operands (the branches of the addition law), byte trees, random points."""
    from oracle import ec as oec
    A = vmx.arithm
    rnd = random.Random(n * 31 + len(curve))
    G, OG = engine_group(vmx, curve), oracle_group(curve)
    R = G.getPRing()
    vals = lambda arr: [elem_value(e) for e in arr.elements()]
    ring = lambda xs: R.toElementArray([R.toElement(x) for x in xs])
    exps = [rnd.randrange(OG.q) for _ in range(n)]
    for i, v in enumerate((0, 1, OG.q - 1, 2)):
        if i < n:
            exps[i] = v
    E = ring(exps)
    X = G.getg().exp(E)                                                        # fixed base
    xs = [OG.op_exp(OG.g, e) for e in exps]
    assert vals(X) == xs
    ys_e = [rnd.randrange(OG.q) for _ in range(n)]
    Y = G.getg().exp(ring(ys_e))
    ys = vals(Y)
    assert vals(X.mul(Y)) == [OG.op_mul(a, b) for a, b in zip(xs, ys)]
    assert vals(X.mul(X)) == [OG.op_mul(a, a) for a in xs]                     # doubling branch
    assert all(v.is_unit() for v in vals(X.mul(X.inv())))                      # opposite points
    assert vals(X.inv()) == [OG.op_inv(a) for a in xs]
    f = [rnd.randrange(OG.q) for _ in range(n)]
    f[0] = 0
    F = ring(f)
    assert vals(X.exp(F)) == [OG.op_exp(a, k) for a, k in zip(xs, f)]          # variable base, per element
    sc = R.toElement(rnd.randrange(1 << 200))
    assert vals(X.exp(sc)) == [OG.op_exp(a, sc.value) for a in xs]             # variable base, one exponent
    # simultaneous scalar multiples sc * X_i + f_i * Y_i (units, equal and opposite operands among them)
    want = [OG.op_mul(OG.op_exp(a, sc.value), OG.op_exp(b, k)) for a, b, k in zip(xs, ys, f)]
    assert vals(X.expMulExp(sc, Y, F)) == want
    assert vals(X.expMulExp(sc, X, F)) == [OG.op_exp(a, (sc.value + k) % OG.q) for a, k in zip(xs, f)]
    assert all(v.is_unit() for v in vals(X.expMulExp(sc, X.inv(), ring([sc.value] * n))))
    assert elem_value(X.expProd(F)) == oar.g_exp_prod(OG, xs, f)               # multi-exponentiation
    short = [rnd.randrange(1 << 100) for _ in range(n)]
    assert elem_value(X.expProd(ring(short))) == oar.g_exp_prod(OG, xs, short)
    assert elem_value(X.prod()) == oar.g_prod(OG, xs)
    pk = A.PPGroup(G, 2).product(X, Y)
    both = pk.expProd(F)
    assert [elem_value(c) for c in both.comps] == [oar.g_exp_prod(OG, xs, f), oar.g_exp_prod(OG, ys, f)]
    perm = list(range(n))
    rnd.shuffle(perm)
    assert vals(X.permute(A.Permutation(perm))) == oar.permute(xs, perm)
    assert vals(X.shiftPush(G.getg())) == [OG.g] + xs[:-1]
    keep = [i % 3 != 1 for i in range(n)]
    assert vals(X.extract(keep)) == [x for x, k in zip(xs, keep) if k]
    assert vals(X.copyOfRange(1, n)) == xs[1:] and elem_value(X.get(n - 1)) == xs[-1]
    assert X.equals(X.copyOfRange(0, n)) and (n < 2 or not X.equals(X.shiftPush(G.getg())))
    cols = G.expProd([X, Y], [5, -3], 3)
    assert vals(cols) == [OG.op_mul(OG.op_exp(a, 5), OG.op_inv(OG.op_exp(b, 3))) for a, b in zip(xs, ys)]
    # byte trees: identical bytes, round trip, rejection of off-curve points and malformed trees
    tb = X.toByteTree().to_bytes()
    assert tb == OG.leaf_array_tree(xs).to_bytes()
    assert G.toElementArray(n, vmx.eio.ByteTreeReader(tb)).equals(X)
    assert G.getg().toByteTree().to_bytes() == OG.leaf_tree(OG.g).to_bytes()
    assert G.getONE().toByteTree().to_bytes() == OG.leaf_tree(oec.UNIT).to_bytes()
    good = next((v for v in xs if not v.is_unit()), OG.g)
    bad_pt = oec.ECPoint(good.x, (good.y + 1) % OG.p)
    for bad in (OG.leaf_array_tree([bad_pt] + xs[1:]).to_bytes(), tb[:-1] + bytes([tb[-1] ^ 1]), tb[:9] + b"\x07" + tb[10:]):
        try:
            G.toElementArray(n, vmx.eio.ByteTreeReader(bad))
            assert False
        except A.ArithmFormatException:
            pass
    # single elements
    a = engine_elem(vmx, G, good)
    assert elem_value(a.exp(sc)) == OG.op_exp(good, sc.value)
    assert elem_value(a.mul(G.getg())) == OG.op_mul(good, OG.g)
    assert elem_value(a.inv()) == OG.op_inv(good) and elem_value(a.mul(a.inv())).is_unit()
    assert elem_value(G.getONE().mul(a)) == good and elem_value(G.getONE().exp(sc)).is_unit()
    assert elem_value(G.toElement(vmx.eio.ByteTreeReader(OG.leaf_tree(good).to_bytes()))) == good
    # random points (generators): same stream, same points, same stream position afterwards
    prg = vmx.crypto.PRGHeuristic()
    prg.setSeed(seed("ec/rs"))
    ors = SeededRandomSource(seed("ec/rs"))
    assert vals(G.randomElementArray(n, prg, 100)) == OG.random_array(n, ors, 100)
    assert prg.getBytes(9) == ors.get_bytes(9)
    assert vals(G.randomElementArray(n, prg, 100)) == OG.random_array(n, ors, 100)      # odd stream offset
    assert vals(R.randomElementArray(n, prg, 100)) == oar.ring_random_array(OG, n, ors, 100)


def mix_parity(vmx, spec, n, k=3, threshold=2, tmpdir=None, width=1, mode="mixing", maxciph=None, light=False,
               gmp=False):
    """`gmp`: the oracle's array operations on its GMP back end (oracle/accel.py, pinned to Python integers in
    tests/test_oracle_accel.py) -- what lets the 2048-bit cases of the GPU suite run in seconds."""
    OG = oracle_group(spec)
    if not gmp:
        return _mix_parity(vmx, spec, OG, n, k, threshold, tmpdir, width, mode, maxciph, light)
    from oracle import accel
    undo = accel.install(OG, accel.cores())
    try:
        return _mix_parity(vmx, spec, OG, n, k, threshold, tmpdir, width, mode, maxciph, light)
    finally:
        undo()


def _mix_parity(vmx, spec, OG, n, k=3, threshold=2, tmpdir=None, width=1, mode="mixing", maxciph=None, light=False):
    """A whole mix (keys, `threshold` shuffles, threshold decryption with proofs) on the engine and on the
    oracle from the same seeds: every file of the proof directory is byte-identical; the engine's vmnv
    (vmnv.MixNetElGamalVerifyFiatShamirSession) and the oracle's accept it, also after a round trip through a
    directory on disk; corrupted files are rejected by both (BASELINE.json config 3 at test size)."""
    vm = importlib.import_module("verificatum-vmn_b200.vmnv")
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G = engine_group(vmx, spec)
    params = mix.SessionParams(pGroupString="mix-%s" % spec)
    oparams = opr.Params(pgroup_string="mix-%s" % spec)
    rs = vmx.crypto.PRGHeuristic()
    rs.setSeed(seed("mix/dealer"))
    M = vm.MixNetElGamal(G, params, k, threshold, rs, width=width)
    irs = vmx.crypto.PRGHeuristic()
    irs.setSeed(seed("mix/input"))
    # the oracle deals the same keys from the same stream, so the input ciphertexts coincide
    probe = SeededRandomSource(seed("mix/dealer"))
    poly0 = oar.ring_random_element(OG, probe, 100)
    opk = (OG.g, OG.op_exp(OG.g, poly0))
    if width == 1:
        w = mix.demoCiphertexts(M.fullPublicKey, n, irs)
        ow = opr.demo_ciphertexts(OG, opk, n, SeededRandomSource(seed("mix/input")))
    else:   # width-omega ciphertexts (elgamal/ProtocolElGamal.java:769-800): widePk^r, r in the product ring
        r = mix.getPlainPGroup(G, width).getPRing().randomElementArray(n, irs, 100)
        w = mix.getWidePublicKey(M.fullPublicKey, width).exp(r)
        ors = SeededRandomSource(seed("mix/input"))
        ow = oar.g_exp(OG, opr.wide_key(opk, width), tuple(oar.ring_random_array(OG, n, ors, 100) for _ in range(width)))
    plain = M.run(w, mode=mode, maxciph=maxciph)
    assert col_values(w) == ow
    od, oplain = opr.run_mix(OG, oparams, k, threshold, ow, SeededRandomSource(seed("mix/dealer")), width=width,
                             mode=mode, maxciph=maxciph)
    assert M.nizkp["type"] == mode.encode()
    assert set(od) == set(M.nizkp)
    for name in sorted(od):
        assert od[name] == M.nizkp[name], name
    assert col_values(plain) == oplain
    # decrypting mixes the plaintexts: same multiset as the inputs' plaintexts m_i (product as a cheap witness)
    V = vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold)
    nizkp = M.nizkp
    if tmpdir is not None:
        nizkp.write(str(tmpdir))
        nizkp = vm.ProofDirectory.read(str(tmpdir))
        assert dict(nizkp) == dict(M.nizkp)
    rep = V.verify(nizkp)
    orep = opr.verify_mix(OG, oparams, k, threshold, dict(nizkp))
    assert rep["accepted"] and orep["accepted"]
    shuffled = mode != "decryption"
    assert rep["shuffles"] == orep["shuffles"] == ({l: True for l in range(1, threshold + 1)} if shuffled else {})
    assert rep["poscs"] == orep["poscs"] == ({l: True for l in range(1, threshold + 1)}
                                             if shuffled and maxciph is not None else {})
    assert rep["decryption"] == orep["decryption"] == (None if mode == "shuffling" else True)
    # the scalar test vectors of `vmnv -t` (global prefix, seeds, challenges, parameters), in the reference's order
    assert rep["vectors"] == orep["vectors"] and any(nm == "der.rho" for nm, _, _ in rep["vectors"])
    if mode != "mixing" or maxciph is not None:
        return _mix_variants(vmx, vm, V, G, OG, params, oparams, k, threshold, M.nizkp, mode, maxciph, light)

    def both_reject(bad):
        for fn, exc in ((lambda: V.verify(bad), vm.VerificationError),
                        (lambda: opr.verify_mix(OG, oparams, k, threshold, dict(bad)), opr.MixVerificationError)):
            try:
                r = fn()
                assert not r["accepted"], "corrupted proof accepted"
            except exc:
                pass

    def flipped(name, pos, bit=1):
        bad = vm.ProofDirectory(M.nizkp)
        raw = bytearray(bad[name])
        raw[pos] ^= bit
        bad[name] = bytes(raw)
        return bad
    both_reject(flipped("proofs/PoSReply01.bt", -3))
    both_reject(flipped("proofs/DecrFactReply02.bt", -1))
    both_reject(flipped("proofs/DecryptionFactors03.bt", -2))
    both_reject(flipped("Plaintexts.bt", -3))
    missing = vm.ProofDirectory(M.nizkp)
    del missing["proofs/PoSCommitment02.bt"]
    both_reject(missing)


def _mix_variants(vmx, vm, V, G, OG, params, oparams, k, threshold, honest, mode, maxciph, light=False):
    """The engine's vmnv and the oracle's on variants of one honest directory of a shuffling / decryption /
    pre-computed session: the options of vmnv (what is verified, the expected type) and one corruption per kind of
    file -- same verdict per party, same accept / reject / fail-stop (MixNetElGamalVerifyFiatShamirSession.java:1318-1668)."""
    def outcome_engine(d, **kw):
        Vk = vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold, **kw) if kw else V
        try:
            r = Vk.verify(d)
            return ("verdict", r["accepted"], r["shuffles"], r["poscs"], r["decryption"], r.get("plaintexts"))
        except vm.VerificationError:
            return ("failstop",)

    def outcome_oracle(d, **kw):
        okw = {{"expectedType": "expected_type"}.get(a, a): b for a, b in kw.items()}
        try:
            r = opr.verify_mix(OG, oparams, k, threshold, dict(d), **okw)
            return ("verdict", r["accepted"], r["shuffles"], r["poscs"], r["decryption"], r.get("plaintexts"))
        except opr.MixVerificationError:
            return ("failstop",)

    options = (dict(dec=False), dict(posc=False), dict(ccpos=False), dict(posc=False, ccpos=False),
               dict(expectedType=mode), dict(expectedType="mixing" if mode != "mixing" else "shuffling"))
    for kw in (options[2:3] + options[5:] if light else options):
        a, b = outcome_engine(honest, **kw), outcome_oracle(honest, **kw)
        assert a == b, (kw, a, b)
    assert outcome_engine(honest, expectedType="mixing" if mode != "mixing" else "shuffling") == ("failstop",)
    seen = set()
    few = ("Reply01.bt", "KeepList01.bt", "Plaintexts.bt", "ShuffledCiphertexts.bt", "DecryptionFactors01.bt")
    for name in sorted(honest):
        kind = "".join(ch for ch in name if not ch.isdigit())
        if kind in seen:       # one file of every kind (party 1's)
            continue
        seen.add(kind)
        if light == "min" and not name.endswith(few):   # (the GPU tests: the variants are host logic, tested on the CPU)
            continue
        raw = bytes(honest[name])
        variants = [raw[:len(raw) // 2]] if light else [raw[:len(raw) // 2], raw + b"\x00"]
        if len(raw) > 8:
            for pos in ((len(raw) - 2,) if light else (len(raw) - 2, 6)):
                b_ = bytearray(raw)
                b_[pos] ^= 0x04
                variants.append(bytes(b_))
        for bad_bytes in variants:
            bad = vm.ProofDirectory(honest)
            bad[name] = bad_bytes
            a, b = outcome_engine(bad), outcome_oracle(bad)
            assert a == b, (name, len(bad_bytes), a, b)
            if len(bad_bytes) == len(raw) and name.endswith(".bt") and "PolynomialInExponent" not in name:
                assert a == ("failstop",) or a[1] is False or not all(a[2].values()) or not all(a[3].values()), (name, a)
        if light and name.endswith(".bt") and not name.endswith(("Commitment01.bt", "Reply01.bt")):
            continue
        missing = vm.ProofDirectory(honest)
        del missing[name]
        a, b = outcome_engine(missing), outcome_oracle(missing)
        assert a == b, (name, "missing", a, b)
    import numpy as np
    flags = np.frombuffer(bytes(honest["proofs/KeepList01.bt"][5:]), dtype=np.uint8).copy() if maxciph is not None else None
    if flags is not None and (flags == 0).any():   # a keep list that keeps the wrong elements (right number): the CCPoS is rejected
        bad = vm.ProofDirectory(honest)
        kl = bytearray(bad["proofs/KeepList01.bt"])
        i, j = int(np.flatnonzero(flags == 1)[0]), int(np.flatnonzero(flags == 0)[0])
        flags[i], flags[j] = 0, 1
        bad["proofs/KeepList01.bt"] = bytes(kl[:5]) + flags.tobytes()
        a, b = outcome_engine(bad), outcome_oracle(bad)
        # (in a mixing session the decryption proof then speaks of another list: fail-stop)
        assert a == b and (a == ("failstop",) if mode == "mixing" else a[2][1] is False and a[1] is False), (a, b)


def ec_edge_cases(vmx, curve):
    """Unit elements inside arrays, zero exponents everywhere, empty arrays, range violations."""
    from oracle import ec as oec
    A = vmx.arithm
    G, OG = engine_group(vmx, curve), oracle_group(curve)
    R = G.getPRing()
    vals = lambda arr: [elem_value(e) for e in arr.elements()]
    ring = lambda xs: R.toElementArray([R.toElement(x) for x in xs])
    pts = [oec.UNIT, OG.g, OG.op_exp(OG.g, 5), oec.UNIT, OG.op_inv(OG.g), OG.op_exp(OG.g, OG.q - 2)]
    X = G.toElementArray([engine_elem(vmx, G, P) for P in pts])
    n = len(pts)
    assert vals(X) == pts
    zeros = ring([0] * n)
    ones = ring([1] * n)
    mixed = [0, 7, OG.q - 1, 3, 1, 2]
    assert vals(X.exp(zeros)) == [oec.UNIT] * n and vals(X.exp(ones)) == pts
    assert vals(X.exp(ring(mixed))) == [OG.op_exp(P, k) for P, k in zip(pts, mixed)]
    assert vals(X.exp(R.toElement(0))) == [oec.UNIT] * n
    assert elem_value(X.expProd(zeros)).is_unit()
    assert elem_value(X.expProd(ring(mixed))) == oar.g_exp_prod(OG, pts, mixed)
    assert elem_value(X.prod()) == oar.g_prod(OG, pts)
    assert vals(X.mul(X.inv())) == [oec.UNIT] * n
    assert vals(X.mul(X)) == [OG.op_mul(P, P) for P in pts]
    rev = G.toElementArray([engine_elem(vmx, G, P) for P in reversed(pts)])
    assert vals(X.mul(rev)) == [OG.op_mul(P, Q) for P, Q in zip(pts, reversed(pts))]
    assert vals(G.getg().exp(zeros)) == [oec.UNIT] * n
    assert vals(G.getONE().exp(ring(mixed))) == [oec.UNIT] * n            # fixed base = unit element
    assert vals(G.expProd([X, rev], [0, -1], 1)) == [OG.op_inv(P) for P in reversed(pts)]
    units = G.toElementArray(4, G.getONE())
    assert vals(units) == [oec.UNIT] * 4 and elem_value(units.prod()).is_unit()
    assert elem_value(units.expProd(ring([5, 6, 7, 8]))).is_unit()
    empty = G.toElementArray([])
    assert empty.size() == 0 and empty.to_matrix().shape == (0, G.elem_bytes)
    assert elem_value(empty.prod()).is_unit() and elem_value(empty.expProd(ring([]))).is_unit()
    assert G.toElementArray(0, vmx.eio.ByteTreeReader(empty.toByteTree().to_bytes())).size() == 0
    # coordinates >= p, a negative coordinate other than the unit's, points off the curve: ArithmFormatException
    cb = G.coord_bytes
    good = OG.g
    for x, y in ((OG.p, good.y), (good.x, OG.p + 1), (-2, -2), (-1, good.y), (good.x, good.y ^ 1)):
        raw = x.to_bytes(cb, "big", signed=True) + y.to_bytes(cb, "big", signed=True)
        try:
            G.toElementArray(1, np.frombuffer(raw, dtype=np.uint8))
            assert False, (x, y)
        except A.ArithmFormatException:
            pass
    try:
        X.mul(units)
        assert False
    except vmx._native.VmxError as e:
        assert e.status == vmx._native.VMX_ESIZE


def _wide_instance(vmx, spec, width, n, label):
    """Width-omega ciphertexts (BASELINE.json config 4: multi-block ciphertexts): w = widePk^{r}, r in the product ring."""
    A = vmx.arithm
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G, OG = engine_group(vmx, spec), oracle_group(spec)
    rs = vmx.crypto.PRGHeuristic()
    rs.setSeed(seed(label + "/setup"))
    x = G.getPRing().randomElement(rs, 100)
    pk = A.PPGroup(G, 2).product(G.getg(), G.getg().exp(x))
    wide = mix.getWidePublicKey(pk, width)
    r = mix.getPlainPGroup(G, width).getPRing().randomElementArray(n, rs, 100)
    w = wide.exp(r)
    ors = SeededRandomSource(seed(label + "/setup"))
    ox = oar.ring_random_element(OG, ors, 100)
    owide = ((OG.g,) * width, (OG.op_exp(OG.g, ox),) * width)
    ow = oar.g_exp(OG, owide, tuple(oar.ring_random_array(OG, n, ors, 100) for _ in range(width)))
    assert col_values(w) == ow
    return G, OG, pk, w, owide, ow


def wide_shuffle_parity(vmx, spec, width, n):
    """Shuffle + PoS of width-omega ciphertexts: identical bytes, cross verification, rejection."""
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G, OG, pk, w, owide, ow = _wide_instance(vmx, spec, width, n, "wide")
    params = mix.SessionParams(pGroupString="wide-%s" % spec)
    oparams = opr.Params(pgroup_string="wide-%s" % spec)
    prs = vmx.crypto.PRGHeuristic()
    prs.setSeed(seed("wide/prover"))
    proof, out = mix.ShufflerSession(G, pk, params, prs).shuffle(width, w, keep_output=True)
    oh = opr.independent_generators(OG, "sha256", oparams.prefix(), "generators", n, 100)
    owp, oproof = opr.shuffle_and_prove(OG, oparams, owide, ow, oh, SeededRandomSource(seed("wide/prover")))
    assert col_values(out) == owp
    assert dataclasses.asdict(proof) == oproof
    verifier = mix.ShufflerSession(G, pk, params, None)
    ok, out2 = verifier.verify(width, w, proof)
    assert ok and col_values(out2) == owp
    assert opr.verify_shuffle(OG, oparams, owide, ow, oh, oproof)
    raw = bytearray(proof.reply)
    raw[-9] ^= 2
    bad = dataclasses.replace(proof, reply=bytes(raw))
    assert verifier.verify(width, w, bad)[0] is False
    assert opr.verify_shuffle(OG, oparams, owide, ow, oh, dataclasses.asdict(bad)) is False


def wide_committed_shuffle_parity(vmx, spec, width, maxciph, n):
    """Pre-computation + commitment-consistent shuffle (CCPoS: pure multi-exponentiation over (1 + 2 omega) N
    elements, the verify of config 4) of width-omega ciphertexts."""
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G, OG, pk, w, owide, ow = _wide_instance(vmx, spec, width, n, "widecc")
    params = mix.SessionParams(pGroupString="widecc-%s" % spec)
    oparams = opr.Params(pgroup_string="widecc-%s" % spec)
    prs = vmx.crypto.PRGHeuristic()
    prs.setSeed(seed("widecc/prover"))
    prover = mix.ShufflerSession(G, pk, params, prs)
    cs = mix.CommittedShuffler(prover, width, maxciph)
    pub = cs.precomp()
    keep = cs.shrink(n)
    prover.randomSource.setSeed(seed("widecc/prove2"))
    proof, out = cs.shuffle(w, keep_output=True)
    oh = opr.independent_generators(OG, "sha256", oparams.prefix(), "generators", maxciph, 100)
    ostate, opub = opr.precomp(OG, oparams, owide, oh, SeededRandomSource(seed("widecc/prover")))
    okeep = opr.shrink(OG, ostate, n)
    owp, oproof = opr.committed_shuffle(OG, oparams, owide, ostate, ow, SeededRandomSource(seed("widecc/prove2")))
    assert pub == opub and keep == okeep and col_values(out) == owp and dataclasses.asdict(proof) == oproof
    verifier = mix.ShufflerSession(G, pk, params, None)
    gens = verifier.deriveGenerators(maxciph)
    pcv = mix.PermutationCommitment(verifier, gens)
    assert pcv.verify(*opub) is True and pcv.shrink(n, okeep) == okeep
    sg = gens.copyOfRange(0, n)
    ok, out2 = mix.verifyCommittedShuffle(verifier, width, sg, pcv.commitment, w, proof)
    assert ok and col_values(out2) == owp
    assert opr.ccpos_verify(OG, oparams, OG.g, oh[:n], ostate["u"], owide, ow, owp, proof.commitment, proof.reply)
    raw = bytearray(proof.reply)
    raw[-2] ^= 8
    okb, outb = mix.verifyCommittedShuffle(verifier, width, sg, pcv.commitment, w, dataclasses.replace(proof, reply=bytes(raw)))
    assert okb is False and col_values(outb) == ow


# ---------------------------------------------------------------------------------------- production-shape kernels
def production_kernels(vmx, bits, n, fixed_windows=(16, 17), mexp_window=12, var_chunk=0):
    """The launches a BASELINE-sized step makes -- thread-per-element k_exp_var / k_exp_var2 (n above the
    cooperative-kernel bound), k_exp_fixed over tables of 16/17-bit windows (split over `parts`), Pippenger at
    c = 12 with length-sorted chunks, k_inv_up/down, the Z_q scans -- bit for bit against the GMP oracle
    (oracle/accel.py, itself pinned to Python integers in tests/test_oracle_accel.py).
    Reference semantics: hvzk/PoSBasicTW.java:1028-1035 (B-chain), elgamal/DistrElGamalSession.java:377-385
    (decryption factors), hvzk/PoSBasicTW.java:407-410,1020-1021 (multi-exponentiations)."""
    from oracle import accel
    rnd = random.Random(bits * 31 + n)
    A = vmx.arithm
    p, q, g = group_params(bits)
    G = A.ModPGroup(p, q, g)
    R = G.getPRing()
    OG = oar.ModPGroup(p, q, g)
    acc = accel.Accel(p, q=q)
    lq = q.bit_length()
    vals = lambda arr: [e.value for e in arr.elements()]
    ring = lambda es: R.toElementArray([A.PFieldElement(R, e) for e in es])

    full = [rnd.randrange(q) for _ in range(n)]
    full[:4] = [0, 1, q - 1, 1 << (lq - 1)]
    xs = acc.exp_fixed(g, [rnd.randrange(q) for _ in range(n)], lq)
    xs[:3] = [1, p - 1, g]
    ys = xs[1:] + xs[:1]
    X = G.toElementArray([A.PGroupElement(G, x) for x in xs])
    Y = G.toElementArray([A.PGroupElement(G, y) for y in ys])
    assert vals(X) == xs
    if var_chunk:
        G.set_tuning(var_chunk=var_chunk)

    # fixed base, full-length exponents, at the windows the cost model picks for N = 10^5 .. 10^6
    E = ring(full)
    want = acc.exp_fixed(g, full, lq)
    for w in fixed_windows:
        G.set_tuning(fixed_window=w)
        assert vals(G.getg().exp(E)) == want, "k_exp_fixed w=%d" % w
    G.set_tuning(fixed_window=0)

    # variable base: per-element 613-bit exponents (B_shift.exp(k_E)), one L_q-bit exponent for all (decryption factors)
    k613 = [rnd.randrange(1 << 613) % q for _ in range(n)]
    k613[:3] = [0, 1, ((1 << 613) - 1) % q]
    K = ring(k613)
    assert vals(X.exp(K)) == acc.exp_var(xs, k613), "k_exp_var per-element"
    s = rnd.randrange(q >> 1, q - (1 << 70))
    assert vals(X.exp(A.PFieldElement(R, s))) == acc.exp_var(xs, s), "k_exp_var common exponent"
    assert vals(X.exp(E)) == acc.exp_var(xs, full), "k_exp_var full-length per-element"

    # simultaneous exponentiation x^v * y^k (the verifier's B-chain check)
    v = rnd.randrange(1 << 256)
    got = vals(X.expMulExp(A.PFieldElement(R, v), Y, K))
    assert got == acc.mul(acc.exp_var(xs, v), acc.exp_var(ys, k613)), "k_exp_var2"

    # multi-exponentiation at the production window
    G.set_tuning(mexp_window=mexp_window)
    e256 = [rnd.randrange(1 << 256) for _ in range(n)]
    e256[:2] = [0, (1 << 256) - 1]
    assert X.expProd(ring(e256)).value == acc.expprod(xs, e256), "Pippenger 256"
    assert X.expProd(K).value == acc.expprod(xs, k613), "Pippenger 613"
    assert X.expProd(E).value == acc.expprod(xs, full), "Pippenger L_q"
    G.set_tuning(mexp_window=0)

    # element-wise product, product of all, inversion by Montgomery's trick
    assert vals(X.mul(Y)) == acc.mul(xs, ys)
    pr = 1
    for x in xs:
        pr = pr * x % p
    assert X.prod().value == pr
    inv = vals(X.inv())
    assert acc.mul(inv, xs) == [1] * n and inv[:3] == [1, p - 1, pow(g, -1, p)]

    # Z_q scans at this size
    a = [rnd.randrange(q) for _ in range(n)]
    b = [rnd.randrange(q) for _ in range(n)]
    Ar, Br = ring(a), ring(b)
    assert vals(Ar.prods()) == oar.r_prods(OG, a)
    x_, d_ = Br.recLin(Ar)
    ox, od = oar.r_rec_lin(OG, b, a)
    assert vals(x_) == ox and d_.value == od
    assert Ar.innerProduct(Br).value == oar.r_inner(OG, a, b)


def with_env(monkeypatch, **env):
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))


def malformed_proof_files(vmx, spec, n, k=3, threshold=2):
    """Every file of a proof directory replaced in turn by (a) thousands of nested node headers, (b) nothing,
    (c) two bytes: the verifier answers with a verdict or VerificationError (the reference's fail-stop), never a
    stray parser exception (RecursionError, EIOException, ValueError), and the oracle agrees on accept/reject."""
    vm = importlib.import_module("verificatum-vmn_b200.vmnv")
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G, OG = engine_group(vmx, spec), oracle_group(spec)
    params = mix.SessionParams(pGroupString="mix-%s" % spec)
    oparams = opr.Params(pgroup_string="mix-%s" % spec)
    rs = vmx.crypto.PRGHeuristic()
    rs.setSeed(seed("mix/dealer"))
    M = vm.MixNetElGamal(G, params, k, threshold, rs)
    irs = vmx.crypto.PRGHeuristic()
    irs.setSeed(seed("mix/input"))
    M.run(mix.demoCiphertexts(M.fullPublicKey, n, irs)).free()
    V = vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold)
    assert V.verify(M.nizkp)["accepted"]
    nested = b"\x00\x00\x00\x00\x01" * 5000
    for name in sorted(M.nizkp):
        if name.endswith(("02.bt", "03.bt")) and "DecrFact" not in name:
            continue      # party 1's file of every kind, all the decryption proof files (the ADVICE cases), header files
        for junk in ((nested, b"", b"\x00\x01") if "DecrFact" in name or "PoSCommitment01" in name else (nested, b"")):
            bad = vm.ProofDirectory(M.nizkp)
            bad[name] = junk
            try:
                accepted = V.verify(bad)["accepted"]
            except vm.VerificationError:
                accepted = False
            try:
                oaccepted = opr.verify_mix(OG, oparams, k, threshold, dict(bad))["accepted"]
            except opr.MixVerificationError:
                oaccepted = False
            assert accepted == oaccepted, (name, len(junk), accepted, oaccepted)
            # the combined check uses the replies of the first `threshold` correct parties only
            # (elgamal/DistrElGamalSessionBasic.java:642-678): a malformed reply of a later party is ignored
            unused = [vm.ProofDirectory.DFRfile(l) for l in range(threshold + 1, k + 1)]
            assert not accepted or name in unused, (name, len(junk))
    # the auxiliary session identifier is read from the proof and enters the global prefix
    # (mixnet/MixNetElGamalVerifyFiatShamirSession.java:160,369-395): another valid one fails every proof, a
    # mismatch with the expected one is fail-stop
    other = vm.ProofDirectory(M.nizkp)
    other["auxsid"] = b"another_session"
    rep = None
    try:
        rep = V.verify(other)
    except vm.VerificationError:
        pass
    assert rep is None or not rep["accepted"]
    try:
        vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold, expectedAuxsid="elsewhere").verify(M.nizkp)
        assert False
    except vm.VerificationError:
        pass
    assert vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold,
                                                   expectedAuxsid=params.auxsid).verify(M.nizkp)["accepted"]
    M2 = vm.MixNetElGamal(G, params, k, threshold, rs, auxsid="second run")
    M2.run(mix.demoCiphertexts(M2.fullPublicKey, n, irs)).free()
    assert V.verify(M2.nizkp)["accepted"]
    assert opr.verify_mix(OG, oparams, k, threshold, dict(M2.nizkp))["accepted"]


def squaring_selftest(vmx, bits, n, iters=3):
    """The dedicated (block-triangular) squaring against repeated multiplication, word for word, on random
    residues and on the edge patterns of its doubling logic (top bits at the 16-word block boundaries, all-ones
    words, p - 1, small values); then against Python integers."""
    import ctypes as C
    rnd = random.Random(bits + n)
    A = vmx.arithm
    p, q, g = group_params(bits)
    G = A.ModPGroup(p, q, g)
    nl = bits // 32
    R = 1 << bits
    Rinv = pow(R, -1, p)
    vals = [rnd.randrange(1, p) for _ in range(n)]
    # the kernels see x * R mod p: choose the Montgomery images, then map back
    mont = [1, p - 1, p - 2, 2, (1 << (bits - 1)) - 1, (1 << (bits - 1)) % p]
    for blk in range(nl // 16):
        mont.append(((1 << (512 * (blk + 1))) - 1) % p)                    # all-ones up to a block boundary
        mont.append((1 << (512 * blk + 511)) % p)                           # only the top bit of block blk
        mont.append(sum(1 << (512 * b + 511) for b in range(blk + 1)) % p)  # top bits of blocks 0..blk
    vals += [m * Rinv % p or 1 for m in mont if 0 < m < p]
    X = G.toElementArray([A.PGroupElement(G, v) for v in vals])
    eq = C.c_int()
    lib = vmx._native.load()
    for it in (1, iters):
        vmx._native.check(lib.vmx_selftest_sqr(X.h, it, C.byref(eq), None))
        assert eq.value == 1, it
    # and the squaring chain of a variable-base exponentiation against Python
    e = A.PFieldElement(G.getPRing(), 1 << 7)
    G.set_tuning(coop_max=0)
    assert [x.value for x in X.exp(e).elements()] == [pow(v, 1 << 7, p) for v in vals]


def concurrent_threads(vmx, bits, n, rounds=6):
    """One context driven from two host threads at once, as the reference does with its export thread
    (hvzk/CCPoSW.java:116-122: the proof is written to file while the protocol goes on) and its optimistic
    next-output thread (mixnet/ShufflerElGamalSession.java:847-856): thread A serialises arrays (device-to-host
    export) while thread B runs array operations and frees temporaries; every result equals the sequential one.
    The C ABI serialises calls per context (one stream, a recursive mutex)."""
    import threading
    rnd = random.Random(bits * 17 + n)
    A, G, R, p, q, g, xs, es, X, E = _arrays(vmx, bits, n, rnd)
    want_exp = oar.g_exp(oar.ModPGroup(p, q, g), xs, es)
    want_bytes = X.toByteTree().to_bytes()
    errors, exported, computed = [], [], []

    def exporter():
        try:
            for _ in range(rounds):
                Y = X.mul(X)                      # a fresh array: its serialisation is not cached
                Z = Y.copyOfRange(0, n)
                exported.append((bytes(X.copyOfRange(0, n).toByteTree().to_bytes()), [e.value for e in Z.elements()][:3]))
                Y.free()
                Z.free()
        except BaseException as e:   # noqa: surfaced below
            errors.append(e)

    def worker():
        try:
            for _ in range(rounds):
                Y = X.exp(E)
                computed.append([e.value for e in Y.elements()])
                Y.free()
                computed.append(X.expProd(E).value)
        except BaseException as e:
            errors.append(e)

    ts = [threading.Thread(target=exporter), threading.Thread(target=worker)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    sq = [x * x % p for x in xs][:3]
    assert all(b == want_bytes and z == sq for b, z in exported) and len(exported) == rounds
    prod = oar.g_exp_prod(oar.ModPGroup(p, q, g), xs, es)
    assert computed[0::2] == [want_exp] * rounds and computed[1::2] == [prod] * rounds


def native_vmnv_parity(vmx, spec, n, k=3, threshold=2, width=1, thorough=True, mode="mixing", maxciph=None,
                       minimal=False):
    """The native universal verifier (csrc/vmnv_native.cpp, include/vmnv.h) against the Python mirror of
    mixnet/MixNetElGamalVerifyFiatShamirSession on the same proof directories: an honest mix, and the same mix with
    one file at a time corrupted / truncated / emptied / missing -- identical verdicts per shuffle, identical
    accept / reject / fail-stop."""
    vm = importlib.import_module("verificatum-vmn_b200.vmnv")
    vn = importlib.import_module("verificatum-vmn_b200.vmnv_native")
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G = engine_group(vmx, spec)
    params = mix.SessionParams(pGroupString="native-%s" % spec, sid="Session_1")
    rs = vmx.crypto.PRGHeuristic()
    rs.setSeed(seed("native/dealer"))
    M = vm.MixNetElGamal(G, params, k, threshold, rs, width=width, auxsid="run 7")
    irs = vmx.crypto.PRGHeuristic()
    irs.setSeed(seed("native/input"))
    if width == 1:
        w = mix.demoCiphertexts(M.fullPublicKey, n, irs)
    else:
        r = mix.getPlainPGroup(G, width).getPRing().randomElementArray(n, irs, 100)
        w = mix.getWidePublicKey(M.fullPublicKey, width).exp(r)
    M.run(w, mode=mode, maxciph=maxciph).free()
    VP = vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold)
    VN = vn.MixNetElGamalVerifyFiatShamirSessionNative(G, params, k, threshold)

    def outcome(V, d):
        try:
            r = V.verify(d)
            return ("verdict", r["type"], r["accepted"], r["shuffles"], r["poscs"], r["decryption"], r.get("plaintexts"))
        except vm.VerificationError:
            return ("failstop", V.report.get("shuffles"), V.report.get("poscs"))

    honest = outcome(VN, M.nizkp)
    assert honest == outcome(VP, M.nizkp) and honest[:3] == ("verdict", mode, True), honest
    assert VN.report["vectors"] == VP.report["vectors"] and len(VN.report["vectors"]) > 10   # `vmnv -t` test vectors
    assert VN.report["hashed_bytes"] > 0 and VN.report["launches"] > 0
    # what is verified (-nodec, -noposc, -noccpos) and the expected type (-mix, -shuffle, -decrypt)
    options = (dict(dec=False), dict(posc=False), dict(ccpos=False), dict(posc=False, ccpos=False),
               dict(expectedType=mode), dict(expectedType="shuffling" if mode == "mixing" else "mixing"))
    for kw in (options[1:2] if minimal else options if thorough else options[2:3] + options[5:]):
        a = outcome(vn.MixNetElGamalVerifyFiatShamirSessionNative(G, params, k, threshold, **kw), M.nizkp)
        b = outcome(vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold, **kw), M.nizkp)
        assert a == b, (kw, a, b)
    # the -auxsid / -width options
    for kw, ok in ((dict(expectedAuxsid="run 7"), True), (dict(expectedAuxsid="other"), False),
                   (dict(expectedWidth=width), True), (dict(expectedWidth=width + 1), False))[:1 if minimal else 4]:
        got = outcome(vn.MixNetElGamalVerifyFiatShamirSessionNative(G, params, k, threshold, **kw), M.nizkp)
        assert (got[0] == "verdict" and got[2]) == ok, (kw, got)
    names = sorted(M.nizkp)
    if minimal:     # (the GPU tests of the pre-computed sessions: the variants are host logic, tested on the CPU)
        names = [nm for nm in names if nm.endswith(("CCPoSReply01.bt", "KeepList02.bt", "PoSCCommitment02.bt", "PoSReply01.bt"))]
    elif thorough:    # every kind of file once (party 1's of each), every header file
        names = [nm for nm in names if not nm.endswith(("02.bt", "03.bt"))]
    else:
        names = [nm for nm in names if nm.endswith(("PoSReply01.bt", "DecrFactCommitment02.bt", "Plaintexts.bt",
                                                    "CCPoSReply01.bt", "KeepList02.bt", "PoSCCommitment02.bt"))]
    nested = b"\x00\x00\x00\x00\x01" * 3000
    for name in names:
        raw = bytes(M.nizkp[name])
        variants = [raw[:len(raw) // 2]]
        if name.endswith(("PoSReply01.bt", "DecrFactCommitment01.bt", "width", "activethreshold", "maxciph", "type",
                          "KeepList01.bt", "CCPoSCommitment01.bt", "PermutationCommitment01.bt")):
            variants += [b"", nested, raw + b"\x00"]
        if len(raw) > 8:
            for pos in ((len(raw) // 2, 6) if name.endswith("PoSCommitment01.bt") else (len(raw) - 2,)):
                b = bytearray(raw)
                b[pos] ^= 0x04
                variants.append(bytes(b))
        for bad_bytes in variants:
            bad = vm.ProofDirectory(M.nizkp)
            bad[name] = bad_bytes
            a, b = outcome(VN, bad), outcome(VP, bad)
            assert a == b, (name, len(bad_bytes), a, b)
        if thorough or name.endswith("Plaintexts.bt"):
            missing = vm.ProofDirectory(M.nizkp)
            del missing[name]
            assert outcome(VN, missing) == outcome(VP, missing), name


def _fuzz_mutation(rng, nizkp, names):
    """One random structure-aware mutation of a proof directory: (description, {name: bytes or None})."""
    name = names[rng.randrange(len(names))]
    raw = bytes(nizkp[name])
    kind = rng.randrange(11)
    if kind == 0 and raw:                                   # flip one bit anywhere
        pos = rng.randrange(len(raw))
        b = bytearray(raw)
        b[pos] ^= 1 << rng.randrange(8)
        return "%s: bit flip at %d" % (name, pos), {name: bytes(b)}
    if kind == 1 and len(raw) >= 5:                         # damage a byte-tree header (tag, count or length)
        heads = [i for i in range(0, len(raw) - 4) if raw[i] in (0, 1) and raw[i + 1] == 0 and raw[i + 2] == 0][:64]
        pos = heads[rng.randrange(len(heads))] if heads else 0
        b = bytearray(raw)
        field = rng.randrange(3)
        if field == 0:
            b[pos] = rng.choice([0, 1, 2, 0xFF])
        else:
            b[pos + 1:pos + 5] = rng.choice([0, 1, 2, 3, 0x7FFFFFFF, 0xFFFFFFFF, len(raw), rng.randrange(1 << 32),
                                             max(0, int.from_bytes(raw[pos + 1:pos + 5], "big") + rng.choice([-1, 1]))
                                             ]).to_bytes(4, "big")
        return "%s: header at %d" % (name, pos), {name: bytes(b)}
    if kind == 2:                                           # truncate anywhere
        cut = rng.randrange(len(raw) + 1)
        return "%s: truncated to %d" % (name, cut), {name: raw[:cut]}
    if kind == 3:                                           # trailing bytes
        extra = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
        return "%s: %d trailing bytes" % (name, len(extra)), {name: raw + extra}
    if kind == 4 and len(raw) > 12:                         # overwrite a run with 0x00 / 0xff / random
        pos = rng.randrange(len(raw) - 4)
        ln = rng.randrange(1, min(64, len(raw) - pos))
        fill = rng.choice([b"\x00", b"\xff", None])
        b = bytearray(raw)
        b[pos:pos + ln] = (fill * ln) if fill else bytes(rng.randrange(256) for _ in range(ln))
        return "%s: run of %d at %d" % (name, ln, pos), {name: bytes(b)}
    if kind == 5:                                           # swap with another file of the directory
        other = names[rng.randrange(len(names))]
        return "%s <-> %s" % (name, other), {name: bytes(nizkp[other]), other: raw}
    if kind == 6:                                           # missing
        return "%s: missing" % name, {name: None}
    if kind == 7:                                           # deep nesting / empty node / empty leaf / empty file
        v = rng.choice([b"", b"\x00\x00\x00\x00\x00", b"\x01\x00\x00\x00\x00", b"\x00\x00\x00\x00\x01" * 5000,
                        b"\x00\xff\xff\xff\xff", b"\x01\xff\xff\xff\xff", b"\x00\x00\x00\x00\x02" + raw + raw])
        return "%s: replaced by %d special bytes" % (name, len(v)), {name: v}
    if kind == 8 and len(raw) > 16:                         # delete or duplicate a slice (shifts everything behind it)
        pos = rng.randrange(len(raw) - 8)
        ln = rng.randrange(1, min(600, len(raw) - pos))
        if rng.randrange(2):
            return "%s: %d bytes deleted at %d" % (name, ln, pos), {name: raw[:pos] + raw[pos + ln:]}
        return "%s: %d bytes doubled at %d" % (name, ln, pos), {name: raw[:pos + ln] + raw[pos:]}
    if kind == 9 and len(raw) > 40:                         # an element set to 0, 1, p - 1 style patterns: last bytes of a leaf
        pos = rng.randrange(len(raw) - 8)
        b = bytearray(raw)
        b[pos:pos + 8] = rng.choice([b"\x00" * 8, b"\x00" * 7 + b"\x01", b"\xff" * 8, b"\x80" + b"\x00" * 7])
        return "%s: word pattern at %d" % (name, pos), {name: bytes(b)}
    # text files: other integers and junk
    v = rng.choice([b"0", b"-1", b"1", b"2", b"3", b"64", b"65", b"1025", b"99999999999", b" 1", b"1\n", b"+1", b"0x1",
                    b"mixing", b"shuffling", b"", b"\xff\xfe", b"1e3"])
    return "%s: text %r" % (name, v), {name: v}


def native_vmnv_fuzz(vmx, spec, n, rounds, seed_label="fuzz", k=3, threshold=2, width=1, log=None, mode="mixing",
                     maxciph=None, oracle=False):
    """Differential fuzzing of the native universal verifier against the Python mirror: `rounds` random
    structure-aware mutations of an honest proof directory (bit flips, damaged headers, truncations, trailing
    bytes, swapped / missing / deeply nested files, junk in the text files); both must reach the same outcome --
    verdict per shuffle, decryption, plaintexts, accept / reject / fail-stop -- and neither may raise anything but
    VerificationError (mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668 never aborts on a bad proof)."""
    import random as _random
    vm = importlib.import_module("verificatum-vmn_b200.vmnv")
    vn = importlib.import_module("verificatum-vmn_b200.vmnv_native")
    mix = importlib.import_module("verificatum-vmn_b200.mixnet")
    G = engine_group(vmx, spec)
    params = mix.SessionParams(pGroupString="fuzz-%s" % spec, sid="Session_1")
    rs = vmx.crypto.PRGHeuristic()
    rs.setSeed(seed("fuzz/dealer"))
    M = vm.MixNetElGamal(G, params, k, threshold, rs, width=width, auxsid="fuzz")
    irs = vmx.crypto.PRGHeuristic()
    irs.setSeed(seed("fuzz/input"))
    if width == 1:
        w = mix.demoCiphertexts(M.fullPublicKey, n, irs)
    else:
        r = mix.getPlainPGroup(G, width).getPRing().randomElementArray(n, irs, 100)
        w = mix.getWidePublicKey(M.fullPublicKey, width).exp(r)
    M.run(w, mode=mode, maxciph=maxciph).free()
    VP = vm.MixNetElGamalVerifyFiatShamirSession(G, params, k, threshold)
    VN = vn.MixNetElGamalVerifyFiatShamirSessionNative(G, params, k, threshold)

    def outcome(V, d):
        try:
            r = V.verify(d)   # (with the seeds and challenges derived on the way: the test vectors of `vmnv -t`)
            return ("verdict", r["accepted"], r["shuffles"], r["poscs"], r["decryption"], r.get("plaintexts"), r["vectors"])
        except vm.VerificationError:
            return ("failstop", V.report.get("shuffles"), V.report.get("poscs"))

    OG = oracle_group(spec) if oracle else None
    oparams = opr.Params(pgroup_string="fuzz-%s" % spec, sid="Session_1")

    def outcome_oracle(d):   # the oracle's vmnv as a third party (`oracle=True`): same outcome, same derived values
        try:
            r = opr.verify_mix(OG, oparams, k, threshold, {nm: bytes(v) for nm, v in d.items()})
            return ("verdict", r["accepted"], r["shuffles"], r["poscs"], r["decryption"], r.get("plaintexts"), r["vectors"])
        except opr.MixVerificationError:
            return ("failstop",)

    rng = _random.Random(seed_label)
    names = sorted(M.nizkp)
    tally = {}
    for i in range(rounds):
        muts = {}
        what = []
        for _ in range(1 if rng.randrange(4) else 2):      # one mutation, sometimes two at once
            d, m = _fuzz_mutation(rng, M.nizkp, names)
            what.append(d)
            muts.update(m)
        bad = vm.ProofDirectory(M.nizkp)
        for nm, v in muts.items():
            if v is None:
                if nm in bad:
                    del bad[nm]
            else:
                bad[nm] = v
        a, b = outcome(VN, bad), outcome(VP, bad)
        assert a == b, (i, what, a, b)
        if oracle:
            c = outcome_oracle(bad)
            assert c == (a if a[0] == "verdict" else a[:1]), (i, what, a, c)
        key = a[0] if a[0] == "failstop" else ("accepted" if a[1] else "rejected")
        tally[key] = tally.get(key, 0) + 1
        if log is not None:
            log("%4d %-9s %s" % (i, key, "; ".join(what)))
    assert outcome(VN, M.nizkp)[:2] == ("verdict", True)   # and the honest directory still verifies afterwards
    return tally
