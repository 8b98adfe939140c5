"""Worker of the multi-process tests of verificatum-vmn_b200/parallel.py (one process per rank).

CPU mode (tests/test_parallel_gloo.py): gloo + the host-emulation build of the C ABI.
GPU mode (tests/test_gpu_parallel.py, launched by torchrun on a multi-GPU box): nccl + the CUDA build.

Every rank builds the SAME seeded instance twice -- once on plain (single-process) arrays, once on
sharded arrays -- and asserts that every array operation and the whole PoSBasicTW / decryption
transcript are bit-identical.
"""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def seed(label):
    return hashlib.sha256(("vmx-par/" + label).encode()).digest()


def main():
    import torch
    import torch.distributed as dist
    mode = sys.argv[1]
    bits = sys.argv[2] if not sys.argv[2].isdigit() else int(sys.argv[2])   # group: bit size or curve name
    n = int(sys.argv[3])
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    if mode == "gpu":
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        device = local
    else:
        dist.init_process_group("gloo")
        device = None
    vmx = importlib.import_module("verificatum-vmn_b200")
    A = vmx.arithm
    par = importlib.import_module("verificatum-vmn_b200.parallel")
    hv = importlib.import_module("verificatum-vmn_b200.hvzk")
    eg = importlib.import_module("verificatum-vmn_b200.elgamal")
    groups = importlib.import_module("verificatum-vmn_b200.groups")
    cr = vmx.crypto
    if isinstance(bits, str):
        G1 = A.ECqPGroup(bits, device=device or 0)
        GS = par.make_curve_group(bits, device)
    else:
        p, q, g = groups.test512() if bits == 512 else groups.rfc3526(bits)
        G1 = A.ModPGroup(p, q, g, device=device or 0)          # plain: the whole array on this rank
        GS = par.make_group(p, q, g, device)                   # sharded over the ranks

    def rs(label):
        r = cr.PRGHeuristic()
        r.setSeed(seed(label))
        return r

    def same(a, b, what):
        ma, mb = a.to_matrix(), b.to_matrix()
        assert ma.shape == mb.shape and np.array_equal(ma, mb), "%s differs on rank %d" % (what, rank)

    # ---- array operations
    for Gx in (G1, GS):
        Gx.t = {}
    vals = {}
    for name, Gx in (("plain", G1), ("shard", GS)):
        R = Gx.getPRing()
        r0 = rs("arrays")
        X = Gx.randomElementArray(n, r0, 100)
        Y = Gx.randomElementArray(n, r0, 100)
        e = R.randomElementArray(n, r0, 100)
        f = R.toElementArray(A.LargeIntegerArray.random(n, 612, r0, R))
        small = R.toElementArray(A.LargeIntegerArray.random(n, 256, r0, R))
        pi = A.Permutation.random(n, r0, 100)
        sc = R.randomElement(r0, 100)
        el = Gx.getg().exp(sc)
        x_rl, d_rl = e.recLin(small)
        vals[name] = dict(
            X=X, e=e, mul=X.mul(Y), fixed=Gx.getg().exp(e), var=X.exp(f), scal=X.exp(sc), perm=X.permute(pi),
            rperm=e.permute(pi), permi=X.permute(pi.inv()), shift=X.shiftPush(el), rshift=e.shiftPush(sc),
            prods=small.prods(), reclin=x_rl, radd=e.add(f), rmul=e.mul(f), rmuladd=e.mulAdd(sc, f),
            cols=Gx.expProd([X, Y], [3, -2], 2),
            extract=X.extract([i % 3 != 1 for i in range(n)]), extract_none=X.extract([False] * n),
            extract_one=X.extract([i == n - 1 for i in range(n)]),
            rng=X.copyOfRange(n // 3, n - n // 4), rrng=e.copyOfRange(min(1, n - 1), n), full=X.copyOfRange(0, n),
            scalars=(X.expProd(small).value, X.prod().value, e.innerProduct(f).value, e.sum().value, small.prod().value,
                     d_rl.value, X.get(0).value, X.get(n - 1).value, e.get(n // 2).value, f.bitLength(),
                     X.equals(X), X.equals(Y), A.expProdMany([X, Y], f)[1].value))
    for k in vals["plain"]:
        if k == "scalars":
            assert vals["plain"][k] == vals["shard"][k], "scalars differ: %r" % (k,)
        else:
            same(vals["plain"][k], vals["shard"][k], k)

    # ---- PoSBasicTW: the sharded prover's transcript equals the single-GPU one; the sharded verifier accepts
    def pos(Gx):
        R = Gx.getPRing()
        r0 = rs("pos/setup")
        x = R.randomElement(r0, 100)
        y = Gx.getg().exp(x)
        pk = A.PPGroup(Gx, 2).product(Gx.getg(), y)
        mix = importlib.import_module("verificatum-vmn_b200.mixnet")
        w = mix.demoCiphertexts(pk, n, r0)
        h = Gx.randomElementArray(n, rs("pos/generators"), 100)
        prs = rs("pos/prover")
        s = R.randomElementArray(n, prs, 100)
        pi = A.Permutation.random(n, prs, 100)
        factors = pk.exp(s)
        reenc = w.mul(factors)
        wp = reenc.permute(pi.inv())
        P = hv.PoSBasicTW(256, 256, 100, cr.PRGHeuristic(), prs)
        P.precompute(Gx.getg(), h, pi)
        P.setInstance(pk, w, wp, s)
        commitment = P.commit(seed("pos/batch")).to_bytes()
        v = int.from_bytes(seed("pos/challenge"), "big")
        reply = P.reply(v).to_bytes()
        u_bytes = P.u.toByteTree().to_bytes()
        V = hv.PoSBasicTW(256, 256, 100, cr.PRGHeuristic(), None)
        V.precompute(Gx.getg(), h)
        V.setInstance(pk, w, wp)
        V.setPermutationCommitment(vmx.eio.ByteTreeReader(u_bytes))
        V.setBatchVector(seed("pos/batch"))
        V.computeAF()
        V.setCommitment(vmx.eio.ByteTreeReader(commitment))
        V.setChallenge(v)
        ok = V.verify(vmx.eio.ByteTreeReader(reply))
        bad = bytearray(reply)
        bad[-5] ^= 2
        V.setBatchVector(seed("pos/batch"))
        rej = V.verify(vmx.eio.ByteTreeReader(bytes(bad)))
        return wp.toByteTree().to_bytes(), u_bytes, commitment, reply, ok, rej

    t1, t2 = pos(G1), pos(GS)
    assert t1[4] is True and t1[5] is False
    assert t1 == t2, "sharded PoS transcript differs from the single-process one on rank %d" % rank

    # ---- decryption factors + batched proof on shards
    def dec(Gx):
        R = Gx.getPRing()
        r0 = rs("dec/setup")
        k, t = 3, 2
        xs = {l: R.randomElement(r0, 100) for l in range(1, k + 1)}
        ys = {l: Gx.getg().exp(xs[l]) for l in xs}
        u = Gx.randomElementArray(n, r0, 100)
        f = {l: eg.decryptionFactors(u, xs[l], k) for l in xs}
        comb = eg.combineDecryptionFactors(f, [False, True, True, True], k, t)
        E = eg.DistrElGamalSessionBasic(1, k, t, 256, 100, cr.PRGHeuristic())
        E.setInstance(Gx.getg(), u, ys, f, xs[1])
        E.setBatchVector(seed("dec/batch"))
        E.batchInput()
        c = E.commit(rs("dec/commit")).to_bytes()
        v = int.from_bytes(seed("dec/challenge"), "big")
        rep = E.reply(v).to_bytes()
        E.batch(1)
        return comb.toByteTree().to_bytes(), c, rep, E.verify(1, v)

    d1, d2 = dec(G1), dec(GS)
    assert d1[3] is True and d1 == d2, "sharded decryption proof differs on rank %d" % rank

    # ---- the public API a mix-server calls: shuffle -> proof bytes -> verify, Fiat-Shamir included
    def fs(Gx):
        import dataclasses
        mix = importlib.import_module("verificatum-vmn_b200.mixnet")
        R = Gx.getPRing()
        r0 = rs("fs/setup")
        y = Gx.getg().exp(R.randomElement(r0, 100))
        pk = A.PPGroup(Gx, 2).product(Gx.getg(), y)
        w = mix.demoCiphertexts(pk, n, r0)
        params = mix.SessionParams(pGroupString="par-test")
        proof, out = mix.ShufflerSession(Gx, pk, params, rs("fs/prover")).shuffle(1, w, keep_output=True)
        verifier = mix.ShufflerSession(Gx, pk, params, None)
        ok, out2 = verifier.verify(1, w, proof)
        raw = bytearray(proof.commitment)
        raw[len(raw) // 2] ^= 1
        rej, out3 = verifier.verify(1, w, dataclasses.replace(proof, commitment=bytes(raw)))
        # online: the verifier hashes the messages as they are published (the root rank alone hashes when sharded)
        ov = verifier.beginVerify(1, w)
        proof2, _ = mix.ShufflerSession(Gx, pk, params, rs("fs/prover")).shuffle(1, w, publish=ov.publish)
        ok3, out4 = ov.finish(proof2)
        ov = verifier.beginVerify(1, w)
        proof3, _ = mix.ShufflerSession(Gx, pk, params, rs("fs/prover")).shuffle(1, w, publish=ov.publish)
        rej3, _ = ov.finish(dataclasses.replace(proof3, reply=bytes(proof3.reply)[:-1] + b"\x00"))
        same_proof = tuple(bytes(x) for x in dataclasses.astuple(proof2)) == tuple(bytes(x) for x in dataclasses.astuple(proof))
        return dataclasses.astuple(proof), ok, out2.equals(out), rej, out3.equals(w), same_proof, ok3, out4.equals(out), rej3

    f1, f2 = fs(G1), fs(GS)
    assert f1[1:] == (True, True, False, True, True, True, True, False), f1[1:]
    assert f1 == f2, "sharded shuffle session differs on rank %d" % rank

    # ---- pre-computation, shrink and commitment-consistent shuffle on shards (BASELINE.json config 4's protocol):
    # extract / copyOfRange move elements between shards (mixnet/PermutationCommitment.java:390-471)
    def committed(Gx):
        import dataclasses
        mix = importlib.import_module("verificatum-vmn_b200.mixnet")
        R = Gx.getPRing()
        r0 = rs("cs/setup")
        y = Gx.getg().exp(R.randomElement(r0, 100))
        pk = A.PPGroup(Gx, 2).product(Gx.getg(), y)
        maxciph = n + 5
        w = mix.demoCiphertexts(pk, n, r0)
        params = mix.SessionParams(pGroupString="par-committed")
        prover = mix.ShufflerSession(Gx, pk, params, rs("cs/prover"))
        cs = mix.CommittedShuffler(prover, 1, maxciph)
        pub = cs.precomp()
        keep = cs.shrink(n)
        proof, out = cs.shuffle(w, keep_output=True)
        verifier = mix.ShufflerSession(Gx, pk, params, None)
        gens = verifier.deriveGenerators(maxciph)
        pcv = mix.PermutationCommitment(verifier, gens)
        okc = pcv.verify(*pub)
        keep2 = pcv.shrink(n, keep)
        ok, out2 = mix.verifyCommittedShuffle(verifier, 1, gens.copyOfRange(0, n), pcv.commitment, w, proof)
        raw = bytearray(proof.reply)
        raw[-2] ^= 8
        rej, _ = mix.verifyCommittedShuffle(verifier, 1, gens.copyOfRange(0, n), pcv.commitment, w,
                                            dataclasses.replace(proof, reply=bytes(raw)))
        # the verifier that follows the bulletin board (its seed hash starts when the output is published; on shards
        # the root rank alone hashes): same proof bytes from a prover with the same randomness, same verdict
        prover2 = mix.ShufflerSession(Gx, pk, params, rs("cs/prover"))
        cs2 = mix.CommittedShuffler(prover2, 1, maxciph)
        cs2.precomp()
        cs2.shrink(n)
        sg = gens.copyOfRange(0, n)
        ov = mix.OnlineCommittedVerification(verifier, 1, sg, pcv.commitment, w)
        proof2, _ = cs2.shuffle(w, publish=ov.publish)
        ok_on, out_on = ov.finish(proof2)
        assert dataclasses.astuple(proof2) == dataclasses.astuple(proof) and ok_on is True and out_on.equals(out)
        return tuple(bytes(b) for b in pub), bytes(keep), dataclasses.astuple(proof), okc, bytes(keep2), ok, \
            out2.equals(out), rej

    c1, c2 = committed(G1), committed(GS)
    assert c1[3] is True and c1[5] is True and c1[6] is True and c1[7] is False, c1[3:]
    assert c1 == c2, "sharded committed shuffle differs from the single-process one on rank %d" % rank

    # ---- a whole 3-party mix and its vmnv-style verification on shards
    def whole_mix(Gx, mode="mixing", maxciph=None):
        vm = importlib.import_module("verificatum-vmn_b200.vmnv")
        mix = importlib.import_module("verificatum-vmn_b200.mixnet")
        params = mix.SessionParams(pGroupString="par-mix")
        M = vm.MixNetElGamal(Gx, params, 3, 2, rs("mix/dealer"))
        w = mix.demoCiphertexts(M.fullPublicKey, n, rs("mix/input"))
        M.run(w, mode=mode, maxciph=maxciph)
        rep = vm.MixNetElGamalVerifyFiatShamirSession(Gx, params, 3, 2).verify(M.nizkp)
        return dict(M.nizkp), rep

    m1, m2 = whole_mix(G1), whole_mix(GS)
    assert m1[1]["accepted"] and m1 == m2, "sharded mix differs from the single-process one on rank %d" % rank
    # ... and a pre-computed session (PoSC over maxciph generators, keep lists, CCPoS over the shrunk commitments:
    # extract / copyOfRange move elements between shards), mixnet/MixNetElGamalVerifyFiatShamirSession.java:1378-1530
    if n >= 2 and not (isinstance(bits, str) and world > 2):
        m1, m2 = whole_mix(G1, maxciph=n + 3), whole_mix(GS, maxciph=n + 3)
        assert m1[1]["accepted"] and m1[1]["poscs"] == {1: True, 2: True} and m1 == m2, \
            "sharded pre-computed mix differs from the single-process one on rank %d" % rank

    stats = (GS.comm.collectives, GS.comm.bytes_exchanged)
    dist.barrier()
    if rank == 0:
        print("PARALLEL OK world=%d group=%s n=%d collectives=%d bytes=%d" % (world, bits, n, stats[0], stats[1]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
