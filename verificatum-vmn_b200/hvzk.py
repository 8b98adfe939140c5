"""Proofs of shuffle on the engine: mirror of `com.verificatum.protocol.hvzk`.

    PoSBasicTW   -- hvzk/PoSBasicTW.java   (Terelius-Wikstrom proof of a shuffle)
    PoSCBasicTW  -- hvzk/PoSCBasicTW.java  (proof of a shuffle of commitments)
    CCPoSBasicW  -- hvzk/CCPoSBasicW.java  (commitment-consistent proof of a shuffle)
    ChallengerRO -- hvzk/ChallengerRO.java (Fiat-Shamir challenges from a random oracle)
    PoSTW / PoSCTW / CCPoSW -- the non-interactive wrappers (hvzk/PoSTW.java:73-260 etc.) with the
                    bulletin board replaced by in-memory byte trees.

Method names, the order of operations, the order in which the prover consumes its random
source and the explicit free() discipline follow the Java line by line (cited per method), so
that the same seeds give the same transcript.  Every array operation is one call into the C
ABI (include/vmx.h); nothing here computes on group elements on the CPU.
"""
from __future__ import annotations

from typing import Optional

from .arithm import (ArithmFormatException, LargeIntegerArray, Permutation, PGroupElementArray, PPGroupElement,
                     PRingElementArray, expProdTogether)
import struct

from .crypto import AsyncDigest, HashfunctionHeuristic, PRGHeuristic, RandomOracle
from .eio import ByteTreeBasic, ByteTreeContainer, ByteTreeLeaf, ByteTreeReader, EIOException


class ProtocolError(RuntimeError):
    pass


class _NullDigest:
    """Digest of a rank that does not hash (sharded runs: the root rank alone hashes, everybody receives the
    result).  The data still passes through here, so whatever produces it (the gather of a sharded array's
    serialisation, a collective) runs on every rank."""

    nbytes = 0

    def update(self, data) -> None:
        pass

    def abandon(self) -> None:
        pass


class ChallengerRO:
    """hvzk/ChallengerRO.java:96-116.

    `comm` (sharded runs, parallel.Comm): Fiat-Shamir hashing is ONE SHA-256 stream per challenge, so in a
    multi-process run the root rank alone computes it and broadcasts the digest (vbitlen / 8 bytes); the other
    ranks feed a null digest.  Hashing the same gigabytes on every rank buys nothing and costs the host W times
    the memory traffic (round 1: 1.81x end to end at 8 GPUs)."""

    def __init__(self, roHashfunction: HashfunctionHeuristic, globalPrefix: bytes, comm=None):
        self.roHashfunction = roHashfunction
        self.globalPrefix = bytes(globalPrefix)
        self.hashed_bytes = 0
        self.comm = comm if (comm is not None and comm.world > 1) else None

    def _hashes(self) -> bool:
        return self.comm is None or self.comm.rank == 0

    def _share(self, out: Optional[bytes], vbitlen: int) -> bytes:
        if self.comm is None:
            return out
        return self.comm.broadcast_bytes(out, (vbitlen + 7) // 8)

    def challenge(self, data: ByteTreeBasic, vbitlen: int, rbitlen: int = 0) -> bytes:
        if not self._hashes():
            data.update(_NullDigest())
            return self._share(None, vbitlen)
        ro = RandomOracle(self.roHashfunction, vbitlen)
        d = ro.getDigest()
        d.update(self.globalPrefix)
        data.update(d)
        self.hashed_bytes += d.nbytes
        return self._share(d.digest(), vbitlen)

    # The same challenge, streamed: `begin` returns a digest that hashes on a worker thread while the caller
    # keeps the GPU busy with work that does not depend on the challenge; the caller writes the byte tree into
    # it piece by piece (node headers included: `node_header(n)`) and calls `finish`.
    def begin(self, vbitlen: int):
        if not self._hashes():
            d = _NullDigest()
            d.vbitlen = vbitlen
            return d
        d = AsyncDigest(RandomOracle(self.roHashfunction, vbitlen).getDigest())
        d.vbitlen = vbitlen
        d.update(self.globalPrefix)
        return d

    def finish(self, d) -> bytes:
        if isinstance(d, _NullDigest):
            return self._share(None, d.vbitlen)
        out = d.digest()
        self.hashed_bytes += d.nbytes
        return self._share(out, d.vbitlen)


def node_header(n: int) -> bytes:
    return struct.pack(">BI", 0, n)


def _to_positive(b: bytes) -> int:
    """LargeInteger.toPositive(byte[])."""
    return int.from_bytes(b, "big")


def _free(*arrays) -> None:
    for a in arrays:
        if a is not None:
            a.free()


# ====================================================================== PoSBasicTW
class PoSBasicTW:
    """hvzk/PoSBasicTW.java:66."""

    def __init__(self, vbitlen: int, ebitlen: int, rbitlen: int, prg: PRGHeuristic, randomSource):
        self.vbitlen, self.ebitlen, self.rbitlen = vbitlen, ebitlen, rbitlen
        self.prg = prg
        self.randomSource = randomSource
        self.u = self.e = self.b = self.B = self.Bp = self.ipe = self.beta = self.epsilon = None
        self.k_B = self.k_E = self.r = self.s = None

    # -- hvzk/PoSBasicTW.java:379-396
    def _precompute_common(self, g, h) -> None:
        self.size = h.size()
        self.pGroup = g.getPGroup()
        self.pRing = self.pGroup.getPRing()
        self.pField = self.pRing.getPField()
        self.g = g
        self.h = h

    # -- :407-410
    def computeAF(self) -> None:
        self.A, self.F = expProdTogether([self.u, self.w], self.e)

    # -- :421-429 / :495-501
    def setInstance(self, pkey: PPGroupElement, w, wp, s=None) -> None:
        self.pkey = pkey
        self.w = w
        self.wp = wp
        self.s = s

    # -- :436-482 (prover) and :379 (verifier: pi is None)
    def precompute(self, g, h, pi: Optional[Permutation] = None, on_u=None) -> None:
        """`on_u(u)` is called as soon as the permutation commitment exists (its hashing can start)."""
        self._precompute_common(g, h)
        if pi is None:
            return
        self.pi = pi
        # u = (h * g^r) permuted                                       :446-452
        self.r = self.pRing.randomElementArray(self.size, self.randomSource, self.rbitlen)
        tmp1 = g.exp(self.r)
        tmp2 = h.mul(tmp1)
        tmp1.free()
        self.u = tmp2.permute(pi)
        tmp2.free()
        if on_u is not None:
            on_u(self.u)
        # randomizers and blinder A' = g^alpha * prod h_i^epsilon_i    :465-481
        self.alpha = self.pRing.randomElement(self.randomSource, self.rbitlen)
        epsilonBitLength = self.ebitlen + self.vbitlen + self.rbitlen
        epsilonIntegers = LargeIntegerArray.random(self.size, epsilonBitLength, self.randomSource, self.pField)
        self.epsilon = self.pField.toElementArray(epsilonIntegers)
        epsilonIntegers.free()
        self.Ap = g.exp(self.alpha).mul(h.expProd(self.epsilon))

    # -- :505-514
    def setPermutationCommitment(self, btr: ByteTreeReader) -> bool:
        """Returns False if the commitment was malformed and replaced by the trivial one."""
        try:
            self.u = self.pGroup.toElementArray(self.h.size(), btr)
            return True
        except ArithmFormatException:
            self.u = self.h.copyOfRange(0, self.h.size())
            return False

    def getPermutationCommitment(self):
        return self.u

    # -- :533-538
    def setBatchVector(self, prgSeed: bytes) -> None:
        self.prg.setSeed(prgSeed)
        lia = LargeIntegerArray.random(self.size, self.ebitlen, self.prg, self.pField)
        self.e = self.pField.unsafeToElementArray(lia)

    # -- :546-700
    def commitIndependent(self) -> None:
        """The part of commit() that does not depend on the batching vector: every value drawn from the random
        source (in the reference's order b, beta, gamma, delta, phi: :571,621,667,678,688) and C', D', F'.  A
        caller that derives the seed by hashing (PoSTW.prove) runs this while the hash is being computed."""
        if self.b is not None:
            return
        g = self.g
        self.b = self.pRing.randomElementArray(self.size, self.randomSource, self.rbitlen)  # :571
        self.beta = self.pRing.randomElementArray(self.size, self.randomSource, self.rbitlen)  # :621
        self.gamma = self.pRing.randomElement(self.randomSource, self.rbitlen)   # :667
        self.delta = self.pRing.randomElement(self.randomSource, self.rbitlen)   # :678
        ciphPRing = self.pkey.project(0).getPGroup().getPRing()        # :687
        self.phi = ciphPRing.randomElement(self.randomSource, self.rbitlen)
        self.Cp = g.exp(self.gamma)
        self.Dp = g.exp(self.delta)
        self.Fp = self.pkey.exp(self.phi.neg()).mul(self.wp.expProd(self.epsilon))  # :690

    def commit(self, prgSeed: bytes, on_B=None) -> ByteTreeBasic:
        """`on_B(B)` is called as soon as the first array of the commitment exists (its hashing can start)."""
        self.setBatchVector(prgSeed)
        self.commitIndependent()
        g, h = self.g, self.h
        piinv = self.pi.inv()
        self.ipe = self.e.permute(piinv)                               # :552-554
        piinv.free()
        h0 = h.get(0)                                                  # :562
        x, self.d = self.b.recLin(self.ipe)                            # :596-598
        y = self.ipe.prods()                                           # :604
        g_exp_x = g.exp(x)                                             # :606
        h0_exp_y = h0.exp(y)                                           # :608
        self.B = g_exp_x.mul(h0_exp_y)                                 # :610
        _free(g_exp_x, h0_exp_y)
        if on_B is not None:
            on_B(self.B)
        xp = x.shiftPush(x.getPRing().getZERO())                       # :637
        yp = y.shiftPush(y.getPRing().getONE())                        # :638
        _free(y, x)
        xp_mul_epsilon = xp.mul(self.epsilon)                          # :642
        beta_add_prod = self.beta.add(xp_mul_epsilon)                  # :643
        g_exp_beta_add_prod = g.exp(beta_add_prod)                     # :644
        yp_mul_epsilon = yp.mul(self.epsilon)                          # :645
        h0_exp_yp_mul_epsilon = h0.exp(yp_mul_epsilon)                 # :646
        self.Bp = g_exp_beta_add_prod.mul(h0_exp_yp_mul_epsilon)       # :648
        _free(h0_exp_yp_mul_epsilon, yp_mul_epsilon, g_exp_beta_add_prod, beta_add_prod, xp_mul_epsilon, yp, xp)
        return ByteTreeContainer(self.B.toByteTree(), self.Ap.toByteTree(), self.Bp.toByteTree(),
                                 self.Cp.toByteTree(), self.Dp.toByteTree(), self.Fp.toByteTree())

    # -- :780-823
    def setCommitment(self, btr: ByteTreeReader) -> ByteTreeBasic:
        ciphPGroup = self.pkey.getPGroup()
        malformed = False
        self.B = self.Bp = None
        try:
            self.B = self.pGroup.toElementArray(self.size, btr.getNextChild())
            self.Ap = self.pGroup.toElement(btr.getNextChild())
            self.Bp = self.pGroup.toElementArray(self.size, btr.getNextChild())
            self.Cp = self.pGroup.toElement(btr.getNextChild())
            self.Dp = self.pGroup.toElement(btr.getNextChild())
            self.Fp = ciphPGroup.toElement(btr.getNextChild())
        except (EIOException, ArithmFormatException):
            malformed = True
        self.commitmentMalformed = malformed
        if malformed:
            _free(self.B, self.Bp)
            one = self.pGroup.getONE()
            self.B = self.pGroup.toElementArray(self.size, one)
            self.Ap = one
            self.Bp = self.pGroup.toElementArray(self.size, one)
            self.Cp = one
            self.Dp = one
            self.Fp = ciphPGroup.getONE()
        return ByteTreeContainer(self.B.toByteTree(), self.Ap.toByteTree(), self.Bp.toByteTree(),
                                 self.Cp.toByteTree(), self.Dp.toByteTree(), self.Fp.toByteTree())

    # -- :838-848
    def setChallenge(self, integerChallenge: int) -> None:
        if not (0 <= integerChallenge and integerChallenge.bit_length() <= self.vbitlen):
            raise ProtocolError("Malformed challenge!")
        self.v = self.pField.toElement(integerChallenge)

    # -- :856-888
    def reply(self, integerChallenge: int) -> ByteTreeBasic:
        self.setChallenge(integerChallenge)
        v = self.v
        a = self.r.innerProduct(self.ipe)
        c = self.r.sum()
        f = self.s.innerProduct(self.e)
        self.k_A = a.mulAdd(v, self.alpha)
        self.k_B = self.b.mulAdd(v, self.beta)
        self.k_C = c.mulAdd(v, self.gamma)
        self.k_D = self.d.mulAdd(v, self.delta)
        self.k_E = self.ipe.mulAdd(v, self.epsilon)
        self.k_F = f.mulAdd(v, self.phi)
        return self.getReply()

    # -- :970-990
    def _parseReplies(self, ciphPRing, btr: ByteTreeReader) -> bool:
        try:
            self.k_A = self.pRing.toElement(btr.getNextChild())
            self.k_B = self.pRing.toElementArray(self.size, btr.getNextChild())
            self.k_C = self.pRing.toElement(btr.getNextChild())
            self.k_D = self.pRing.toElement(btr.getNextChild())
            self.k_E = self.pField.toElementArray(self.size, btr.getNextChild())
            self.k_F = ciphPRing.toElement(btr.getNextChild())
            return True
        except (EIOException, ArithmFormatException):
            return False

    # -- :1000-1066
    def verify(self, btr: ByteTreeReader) -> bool:
        ciphPRing = self.pkey.project(0).getPGroup().getPRing()
        if not self._parseReplies(ciphPRing, btr):
            return False
        return self.verifyParsed()

    def verifyIndependent(self) -> None:
        """The operands of the five checks that depend on the proof alone, not on the batching vector or the
        challenge: g^k_A * prod h^k_E, g^k_B * B_shift^k_E (the largest item of a verification), g^k_C, g^k_D,
        pk^-k_F * prod w'^k_E and C (:1013,1021,1030-1033,1048,1055,1063).  A verifier that derives its challenges
        by hashing (PoSTW.verify) queues these while the hash is being computed; the verdict is the same
        conjunction of the same five equalities."""
        g, h, u = self.g, self.h, self.u
        h0 = h.get(0)
        ind = {}
        ind["C"] = u.prod().div(h.prod())                                        # :1013
        h_k_E, wp_k_E = expProdTogether([h, self.wp], self.k_E)                  # :1021,1063
        ind["rightA"] = g.exp(self.k_A).mul(h_k_E)                               # :1021
        # B^v * B' == g^k_B * B_shift^k_E (:1028-1035) is checked as B^v * (B_shift^-1)^k_E * B' == g^k_B: the two
        # variable-base exponentiations then share one chain of squarings (expMulExp); the inverses cost three
        # multiplications per element and do not depend on the challenge.
        ind["rightB"] = g.exp(self.k_B)                                          # :1030
        B_shift = self.B.shiftPush(h0)                                           # :1031
        ind["B_shift_inv"] = B_shift.inv()
        _free(B_shift)
        ind["rightC"] = g.exp(self.k_C)                                          # :1048
        ind["rightD"] = g.exp(self.k_D)                                          # :1055
        ind["rightF"] = self.pkey.exp(self.k_F.neg()).mul(wp_k_E)                # :1063
        self._ind = ind

    def verifyParsed(self) -> bool:
        """The five checks of verify() on already-imported replies (:1008-1066)."""
        if getattr(self, "_ind", None) is None:
            self.verifyIndependent()
        ind, self._ind = self._ind, None
        h, v = self.h, self.v
        h0 = h.get(0)
        self.C = ind["C"]
        self.D = self.B.get(self.size - 1).div(h0.exp(self.e.prod()))            # :1014
        verdictA = self.A.expMul(v, self.Ap).equals(ind["rightA"])               # :1020-1021
        both = self.B.expMulExp(v, ind["B_shift_inv"], self.k_E)                 # :1028,1032
        leftSide = both.mul(self.Bp)
        rightSide = ind["rightB"]
        verdictB = leftSide.equals(rightSide)                                    # :1035
        _free(both, leftSide, rightSide, ind["B_shift_inv"])
        verdictC = self.C.expMul(v, self.Cp).equals(ind["rightC"])               # :1048
        verdictD = self.D.expMul(v, self.Dp).equals(ind["rightD"])               # :1055
        verdictF = self.F.expMul(v, self.Fp).equals(ind["rightF"])               # :1062-1063
        self.verdicts = (verdictA, verdictB, verdictC, verdictD, verdictF)
        return verdictA and verdictB and verdictC and verdictD and verdictF

    # -- :1073-1080
    def getReply(self) -> ByteTreeBasic:
        return ByteTreeContainer(self.k_A.toByteTree(), self.k_B.toByteTree(), self.k_C.toByteTree(),
                                 self.k_D.toByteTree(), self.k_E.toByteTree(), self.k_F.toByteTree())

    # -- :1088-1101
    def free(self) -> None:
        ind = getattr(self, "_ind", None)
        if ind is not None:
            _free(ind["rightB"], ind["B_shift_inv"])
            self._ind = None
        _free(self.r, self.u, self.e, self.b, self.B, self.Bp, self.ipe, self.beta, self.epsilon, self.k_B, self.k_E)
        self.r = self.u = self.e = self.b = self.B = self.Bp = self.ipe = self.beta = self.epsilon = None
        self.k_B = self.k_E = None


# ====================================================================== PoSCBasicTW
class PoSCBasicTW:
    """hvzk/PoSCBasicTW.java:65 -- PoSBasicTW without the ciphertext (F) part."""

    def __init__(self, vbitlen, ebitlen, rbitlen, prg, randomSource):
        self.vbitlen, self.ebitlen, self.rbitlen = vbitlen, ebitlen, rbitlen
        self.prg = prg
        self.randomSource = randomSource
        self.e = self.b = self.B = self.Bp = self.ipe = self.beta = self.epsilon = self.k_B = self.k_E = None

    # -- :306-339
    def setInstance(self, g, h, u, r=None, pi=None) -> None:
        self.g, self.h, self.u, self.r, self.pi = g, h, u, r, pi
        self.size = h.size()
        self.pGroup = g.getPGroup()
        self.pRing = self.pGroup.getPRing()
        self.pField = self.pRing.getPField()

    # -- :350-355
    def setBatchVector(self, prgSeed: bytes) -> None:
        self.prg.setSeed(prgSeed)
        lia = LargeIntegerArray.random(self.size, self.ebitlen, self.prg, self.pField)
        self.e = self.pField.unsafeToElementArray(lia)

    # -- :363-529
    def commit(self, prgSeed: bytes) -> ByteTreeBasic:
        self.setBatchVector(prgSeed)
        g, h = self.g, self.h
        piinv = self.pi.inv()
        self.ipe = self.e.permute(piinv)
        h0 = h.get(0)
        self.b = self.pRing.randomElementArray(self.size, self.randomSource, self.rbitlen)
        x, self.d = self.b.recLin(self.ipe)
        y = self.ipe.prods()
        g_exp_x = g.exp(x)
        h0_exp_y = h0.exp(y)
        self.B = g_exp_x.mul(h0_exp_y)
        _free(g_exp_x, h0_exp_y)
        self.alpha = self.pRing.randomElement(self.randomSource, self.rbitlen)
        epsilonBitLength = self.ebitlen + self.vbitlen + self.rbitlen
        epsilonIntegers = LargeIntegerArray.random(self.size, epsilonBitLength, self.randomSource, self.pField)
        self.epsilon = self.pField.toElementArray(epsilonIntegers)
        epsilonIntegers.free()
        self.Ap = g.exp(self.alpha).mul(h.expProd(self.epsilon))
        self.beta = self.pRing.randomElementArray(self.size, self.randomSource, self.rbitlen)
        xp = x.shiftPush(x.getPRing().getZERO())
        yp = y.shiftPush(y.getPRing().getONE())
        _free(y, x)
        xp_mul_epsilon = xp.mul(self.epsilon)
        beta_add_prod = self.beta.add(xp_mul_epsilon)
        g_exp_beta_add_prod = g.exp(beta_add_prod)
        yp_mul_epsilon = yp.mul(self.epsilon)
        h0_exp_yp_mul_epsilon = h0.exp(yp_mul_epsilon)
        self.Bp = g_exp_beta_add_prod.mul(h0_exp_yp_mul_epsilon)
        _free(h0_exp_yp_mul_epsilon, yp_mul_epsilon, g_exp_beta_add_prod, beta_add_prod, xp_mul_epsilon, yp, xp)
        self.gamma = self.pRing.randomElement(self.randomSource, self.rbitlen)
        self.Cp = g.exp(self.gamma)
        self.delta = self.pRing.randomElement(self.randomSource, self.rbitlen)
        self.Dp = g.exp(self.delta)
        return ByteTreeContainer(self.B.toByteTree(), self.Ap.toByteTree(), self.Bp.toByteTree(),
                                 self.Cp.toByteTree(), self.Dp.toByteTree())

    # -- :540-575
    def setCommitment(self, btr: ByteTreeReader) -> ByteTreeBasic:
        malformed = False
        self.B = self.Bp = None
        try:
            self.B = self.pGroup.toElementArray(self.size, btr.getNextChild())
            self.Ap = self.pGroup.toElement(btr.getNextChild())
            self.Bp = self.pGroup.toElementArray(self.size, btr.getNextChild())
            self.Cp = self.pGroup.toElement(btr.getNextChild())
            self.Dp = self.pGroup.toElement(btr.getNextChild())
        except (EIOException, ArithmFormatException):
            malformed = True
        if malformed:
            _free(self.B, self.Bp)
            one = self.pGroup.getONE()
            self.B = self.pGroup.toElementArray(self.size, one)
            self.Ap = one
            self.Bp = self.pGroup.toElementArray(self.size, one)
            self.Cp = one
            self.Dp = one
        return ByteTreeContainer(self.B.toByteTree(), self.Ap.toByteTree(), self.Bp.toByteTree(),
                                 self.Cp.toByteTree(), self.Dp.toByteTree())

    # -- :590-600
    def setChallenge(self, integerChallenge: int) -> None:
        if not (0 <= integerChallenge and integerChallenge.bit_length() <= self.vbitlen):
            raise ProtocolError("Malformed challenge!")
        self.v = self.pField.toElement(integerChallenge)

    # -- :607-636
    def reply(self, integerChallenge: int) -> ByteTreeBasic:
        self.setChallenge(integerChallenge)
        v = self.v
        a = self.r.innerProduct(self.ipe)
        c = self.r.sum()
        self.k_A = a.mulAdd(v, self.alpha)
        self.k_B = self.b.mulAdd(v, self.beta)
        self.k_C = c.mulAdd(v, self.gamma)
        self.k_D = self.d.mulAdd(v, self.delta)
        self.k_E = self.ipe.mulAdd(v, self.epsilon)
        return ByteTreeContainer(self.k_A.toByteTree(), self.k_B.toByteTree(), self.k_C.toByteTree(),
                                 self.k_D.toByteTree(), self.k_E.toByteTree())

    # -- :646-727
    def parseReply(self, btr: ByteTreeReader) -> bool:
        self._ind = None
        try:
            self.k_A = self.pRing.toElement(btr.getNextChild())
            self.k_B = self.pRing.toElementArray(self.size, btr.getNextChild())
            self.k_C = self.pRing.toElement(btr.getNextChild())
            self.k_D = self.pRing.toElement(btr.getNextChild())
            self.k_E = self.pField.toElementArray(self.size, btr.getNextChild())
        except (EIOException, ArithmFormatException):
            return False
        return True

    def verifyIndependent(self) -> None:
        """What the four checks need from the proof alone (cf. PoSBasicTW.verifyIndependent): queued by
        PoSCTW.verify while the seed of the batching vector is being hashed."""
        g, h, u = self.g, self.h, self.u
        h0 = h.get(0)
        ind = {}
        ind["C"] = u.prod().div(h.prod())
        ind["rightA"] = g.exp(self.k_A).mul(h.expProd(self.k_E))
        ind["rightB"] = g.exp(self.k_B)
        B_shift = self.B.shiftPush(h0)
        ind["B_shift_inv"] = B_shift.inv()
        _free(B_shift)
        ind["rightC"] = g.exp(self.k_C)
        ind["rightD"] = g.exp(self.k_D)
        self._ind = ind

    def verifyParsed(self) -> bool:
        if getattr(self, "_ind", None) is None:
            self.verifyIndependent()
        ind, self._ind = self._ind, None
        h, u, v = self.h, self.u, self.v
        h0 = h.get(0)
        A = u.expProd(self.e)
        D = self.B.get(self.size - 1).div(h0.exp(self.e.prod()))
        verdict = A.expMul(v, self.Ap).equals(ind["rightA"])
        # as in PoSBasicTW.verifyParsed: B^v * (B_shift^-1)^k_E * B' == g^k_B, one chain of squarings
        both = self.B.expMulExp(v, ind["B_shift_inv"], self.k_E)
        leftSide = both.mul(self.Bp)
        B_res = leftSide.equals(ind["rightB"])
        _free(both, leftSide, ind["rightB"], ind["B_shift_inv"])
        verdict = verdict and B_res
        verdict = verdict and ind["C"].expMul(v, self.Cp).equals(ind["rightC"])
        verdict = verdict and D.expMul(v, self.Dp).equals(ind["rightD"])
        return verdict

    def verify(self, btr: ByteTreeReader) -> bool:
        if not self.parseReply(btr):
            return False
        return self.verifyParsed()

    def free(self) -> None:
        ind = getattr(self, "_ind", None)
        if ind is not None:
            _free(ind["rightB"], ind["B_shift_inv"])
            self._ind = None
        _free(self.e, self.b, self.B, self.Bp, self.ipe, self.beta, self.epsilon, self.k_B, self.k_E)
        self.e = self.b = self.B = self.Bp = self.ipe = self.beta = self.epsilon = self.k_B = self.k_E = None


# ====================================================================== CCPoSBasicW
class CCPoSBasicW:
    """hvzk/CCPoSBasicW.java:65 (the raisedu/raisedh variant of computeAB/verify, :502-504 and
    :571-579, mixes a basic array into a product-group array inside VCR and is not mirrored)."""

    def __init__(self, vbitlen, ebitlen, rbitlen, prg):
        self.vbitlen, self.ebitlen, self.rbitlen = vbitlen, ebitlen, rbitlen
        self.prg = prg
        self.e = self.ipe = self.epsilon = self.k_E = None

    # -- :268-313
    def setInstance(self, g, h, u, pkey, w, wp, r=None, pi=None, s=None) -> None:
        self.g, self.h, self.u, self.pkey, self.w, self.wp = g, h, u, pkey, w, wp
        self.r, self.pi, self.s = r, pi, s
        self.size = h.size()
        self.pGroup = g.getPGroup()
        self.pRing = self.pGroup.getPRing()
        self.pField = self.pRing.getPField()

    # -- :330-335
    def setBatchVector(self, prgSeed: bytes) -> None:
        self.prg.setSeed(prgSeed)
        lia = LargeIntegerArray.random(self.size, self.ebitlen, self.prg, self.pField)
        self.e = self.pField.unsafeToElementArray(lia)

    # -- :344-396
    def commitIndependent(self, randomSource) -> None:
        """A' and B' (:365-393): functions of the prover's randomness alone (alpha, epsilon, beta, drawn in the
        reference's order) -- the two multi-exponentiations that are nearly all of a CCPoS proof.  CCPoSW.prove
        queues them while the seed of the batching vector is being hashed."""
        self.alpha = self.pRing.randomElement(randomSource, self.rbitlen)
        epsilonBitLength = self.ebitlen + self.vbitlen + self.rbitlen
        epsilonIntegers = LargeIntegerArray.random(self.size, epsilonBitLength, randomSource, self.pField)
        self.epsilon = self.pField.toElementArray(epsilonIntegers)
        epsilonIntegers.free()
        h_eps, wp_eps = expProdTogether([self.h, self.wp], self.epsilon)
        self.Ap = self.g.exp(self.alpha).mul(h_eps)
        ciphPRing = self.pkey.project(0).getPGroup().getPRing()
        self.beta = ciphPRing.randomElement(randomSource, self.rbitlen)
        self.Bp = self.pkey.exp(self.beta.neg()).mul(wp_eps)

    def commit(self, prgSeed: bytes, randomSource) -> ByteTreeBasic:
        self.setBatchVector(prgSeed)
        piinv = self.pi.inv()
        self.ipe = self.e.permute(piinv)
        if self.epsilon is None:
            self.commitIndependent(randomSource)
        return ByteTreeContainer(self.Ap.toByteTree(), self.Bp.toByteTree())

    # -- :406-428
    def setCommitment(self, btr: ByteTreeReader) -> ByteTreeBasic:
        ciphPGroup = self.pkey.getPGroup()
        try:
            self.Ap = self.pGroup.toElement(btr.getNextChild())
            self.Bp = ciphPGroup.toElement(btr.getNextChild())
        except (EIOException, ArithmFormatException):
            self.Ap = self.pGroup.getONE()
            self.Bp = ciphPGroup.getONE()
        return ByteTreeContainer(self.Ap.toByteTree(), self.Bp.toByteTree())

    def setChallenge(self, integerChallenge: int) -> None:
        if not (0 <= integerChallenge and integerChallenge.bit_length() <= self.vbitlen):
            raise ProtocolError("Malformed challenge!")
        self.v = self.pField.toElement(integerChallenge)

    # -- :462-485
    def reply(self, integerChallenge: int) -> ByteTreeBasic:
        self.setChallenge(integerChallenge)
        a = self.r.innerProduct(self.ipe)
        b = self.s.innerProduct(self.e)
        self.k_A = a.mulAdd(self.v, self.alpha)
        self.k_B = b.mulAdd(self.v, self.beta)
        self.k_E = self.ipe.mulAdd(self.v, self.epsilon)
        return ByteTreeContainer(self.k_A.toByteTree(), self.k_B.toByteTree(), self.k_E.toByteTree())

    # -- :493-506
    def computeAB(self) -> None:
        self.A, self.B = expProdTogether([self.u, self.w], self.e)

    # -- :519-584
    def parseReply(self, btr: ByteTreeReader) -> bool:
        ciphPRing = self.pkey.project(0).getPGroup().getPRing()
        self._ind = None
        try:
            self.k_A = self.pRing.toElement(btr.getNextChild())
            self.k_B = ciphPRing.toElement(btr.getNextChild())
            self.k_E = self.pField.toElementArray(self.size, btr.getNextChild())
        except (EIOException, ArithmFormatException):
            self.k_A = self.pRing.getZERO()
            self.k_B = None
            self.k_E = self.pField.toElementArray(self.size, self.pField.getZERO())
            return False
        return True

    def verifyIndependent(self) -> None:
        """The right-hand sides g^k_A * prod h^k_E and pk^-k_B * prod w'^k_E (:554-579): functions of the proof
        alone (the two multi-exponentiations are nearly all of a CCPoS verification), queued by CCPoSW.verify while
        the seed of the batching vector is being hashed."""
        h_k_E, wp_k_E = expProdTogether([self.h, self.wp], self.k_E)
        self._ind = (self.g.exp(self.k_A).mul(h_k_E), self.pkey.exp(self.k_B.neg()).mul(wp_k_E))

    def verifyParsed(self) -> bool:
        if getattr(self, "_ind", None) is None:
            self.verifyIndependent()
        (rightA, rightB), self._ind = self._ind, None
        verdict = True
        if not self.A.expMul(self.v, self.Ap).equals(rightA):
            verdict = False
        if verdict and not self.B.expMul(self.v, self.Bp).equals(rightB):
            verdict = False
        return verdict

    def verify(self, btr: ByteTreeReader) -> bool:
        if not self.parseReply(btr):
            return False
        return self.verifyParsed()

    def free(self) -> None:
        _free(self.e, self.ipe, self.epsilon, self.k_E)
        self.e = self.ipe = self.epsilon = self.k_E = None


# ====================================================================== Fiat-Shamir wrappers
class PoSTW:
    """hvzk/PoSTW.java:52 with the bulletin board replaced by byte strings: `prove` returns the
    three published messages, `verify` consumes them (what vmnv reads from the proof directory:
    PermutationCommitment, PoSCommitment, PoSReply; hvzk/PoSTW.java:281-307)."""

    def __init__(self, vbitlen, ebitlen, rbitlen, prg: PRGHeuristic, randomSource, challenger: ChallengerRO):
        self.vbitlen, self.ebitlen, self.rbitlen = vbitlen, ebitlen, rbitlen
        self.prg, self.randomSource, self.challenger = prg, randomSource, challenger
        self.P = self.V = None
        self._seedDigest = None
        self._seedFed = 0       # how far the streamed seed hash has got: 0 (g, h), 1 (.. u), 2 (.. pk, w)

    # -- :80-88 / :167-173
    def precompute(self, g, h, pi: Optional[Permutation] = None) -> None:
        basic = PoSBasicTW(self.vbitlen, self.ebitlen, self.rbitlen, self.prg, self.randomSource)
        # The seed of the batching vector is RO(g, h, u, pk, w, w') (:118-124): g and h are known now, so a
        # prover starts hashing them BEFORE queueing its exponentiations and the worker thread hashes beside them.
        if pi is not None and getattr(self, "_seedDigest", None) is None:
            self._seedDigest = self._seed_begin(g, h)
        if pi is None:
            basic.precompute(g, h)
            self.V = basic
        else:
            basic.precompute(g, h, pi, on_u=self._seed_u)   # u is hashed while A' is being computed
            self.P = basic

    def _seed_u(self, u) -> None:
        u.toByteTree().update(self._seedDigest)
        self._seedFed = 1

    def continueSeed(self, pkey, w) -> None:
        """Prover: the public key and the input ciphertexts are known before the output exists -- hash them
        (after g, h, u) while the device re-encrypts."""
        if self._seedDigest is not None and self._seedFed == 1:
            pkey.toByteTree().update(self._seedDigest)
            w.toByteTree().update(self._seedDigest)
            self._seedFed = 2

    def beginSeed(self, g, h) -> None:
        """Prover: start hashing (g, h) now -- call it before any device work is queued (the serialisation of
        h is a device-to-host copy that would otherwise wait behind that work)."""
        self._seedDigest = self._seed_begin(g, h)

    def _seed_begin(self, g, h) -> AsyncDigest:
        d = self.challenger.begin(8 * self.prg.minNoSeedBytes())
        d.update(node_header(6))
        g.toByteTree().update(d)
        h.toByteTree().update(d)
        return d

    def _seed_finish(self, d: AsyncDigest, u, pkey, w, wp, fed: int = 0) -> AsyncDigest:
        if fed < 1:
            u.toByteTree().update(d)
        if fed < 2:
            pkey.toByteTree().update(d)
            w.toByteTree().update(d)
        wp.toByteTree().update(d)
        return d

    def _seed(self, B: PoSBasicTW, pkey, w, wp) -> bytes:
        challengeData = ByteTreeContainer(B.g.toByteTree(), B.h.toByteTree(), B.u.toByteTree(), pkey.toByteTree(),
                                          w.toByteTree(), wp.toByteTree())                      # :118-124
        return self.challenger.challenge(challengeData, 8 * self.prg.minNoSeedBytes(), self.rbitlen)

    # -- :95-165
    def prove(self, pkey, w, wp, s, publish=None):
        """`publish(name, message)`: called as each message goes to the bulletin board (:105, :148, :160), so
        that an online verifier (the other mix-servers of the reference wait on the board message by message,
        :195-245) can hash and import it while this prover is still computing."""
        P = self.P
        P.setInstance(pkey, w, wp, s)
        permutationCommitment = P.u.toByteTree().to_buffer()
        if publish is not None:
            publish("permutationCommitment", permutationCommitment)
        d = self._seedDigest if self._seedDigest is not None else self._seed_begin(P.g, P.h)
        fed, self._seedDigest, self._seedFed = self._seedFed, None, 0
        self._seed_finish(d, P.u, pkey, w, wp, fed)
        P.commitIndependent()            # the GPU works on C', D', F' while the worker thread hashes
        prgSeed = self.challenger.finish(d)
        # challenge = RO(node(leaf(seed), commitment)) (:146-147): B is hashed while B' is being computed
        cd = self.challenger.begin(self.vbitlen)
        cd.update(node_header(2))
        ByteTreeLeaf(prgSeed).update(cd)
        cd.update(node_header(6))
        commitment = P.commit(prgSeed, on_B=lambda B: B.toByteTree().update(cd))
        # the message is assembled BEFORE its remaining children are hashed: B' is then serialised by the engine
        # straight into the message (ByteTreeDeviceArray.update / eio._Writer.reserve) and hashed from there
        commitmentBytes = commitment.to_buffer()
        if publish is not None:
            publish("commitment", commitmentBytes)
        for child in commitment.children[1:]:
            child.update(cd)
        challengeBytes = self.challenger.finish(cd)
        reply = P.reply(_to_positive(challengeBytes))
        replyBytes = reply.to_buffer()
        if publish is not None:
            publish("reply", replyBytes)
        out = (permutationCommitment, commitmentBytes, replyBytes)
        P.free()
        return out

    # -- verifier, online: the hashing of a published message starts when it appears on the board, under the
    # premise that it is well formed (then its bytes ARE the byte tree of the parsed value); `verify` checks the
    # premise and hashes again if it does not hold, so verdicts are those of the offline order.
    def prehashSeed(self, pkey, w, permutationCommitment, output) -> None:
        V = self.V
        d = self._seed_begin(V.g, V.h)
        d.update(permutationCommitment)
        pkey.toByteTree().update(d)
        w.toByteTree().update(d)
        d.update(output)
        self._preSeed = (d, permutationCommitment, output)
        self._preChallenge = None

    def prehashChallenge(self, commitment) -> None:
        """Needs the verifier's own seed: waits for the streamed seed hash, then streams the challenge hash."""
        pre = getattr(self, "_preSeed", None)
        if pre is None or isinstance(pre[0], bytes):
            return
        prgSeed = self.challenger.finish(pre[0])
        self._preSeed = (prgSeed, pre[1], pre[2])
        cd = self.challenger.begin(self.vbitlen)
        cd.update(node_header(2))
        ByteTreeLeaf(prgSeed).update(cd)
        cd.update(commitment)
        self._preChallenge = (cd, commitment)

    def _abandon_prehash(self) -> None:
        """Drop streamed hashes whose premise failed (every rank of a sharded run decides alike: the premise is a
        function of the published bytes)."""
        for st in (getattr(self, "_preSeed", None), getattr(self, "_preChallenge", None)):
            if st is not None and not isinstance(st[0], bytes):
                st[0].abandon()
        self._preSeed = self._preChallenge = None

    # -- :177-260
    def verify(self, pkey, w, wp, permutationCommitment: bytes, commitment: bytes, reply: bytes,
               outputBytes=None) -> bool:
        """`outputBytes`: the published message `wp` was parsed from (only needed to validate a streamed seed hash:
        the hash stands only if it covered these very bytes)."""
        V = self.V
        V.setInstance(pkey, w, wp)
        u_parsed = True
        try:
            u_parsed = V.setPermutationCommitment(ByteTreeReader(permutationCommitment)) is not False
        except EIOException:
            V.u = V.h.copyOfRange(0, V.h.size())
            u_parsed = False
        # online verification (prehashSeed): the seed hash was started from the published bytes; it stands if
        # they are exactly the byte trees of the parsed arrays (well formed, no trailing bytes)
        self.u_parsed = u_parsed   # (vmnv reads the commitment itself first and fail-stops when it is malformed)
        pre, prc = getattr(self, "_preSeed", None), getattr(self, "_preChallenge", None)
        self._preSeed = self._preChallenge = None
        pre_ok = pre is not None and u_parsed and pre[1] is permutationCommitment and pre[2] is outputBytes and \
            len(permutationCommitment) == V.u.toByteTree().total_bytes() and len(pre[2]) == wp.toByteTree().total_bytes()
        if pre is not None and not pre_ok:
            self._preSeed, self._preChallenge = pre, prc
            self._abandon_prehash()
            pre = prc = None
        # every input of the seed is on the host already: hash it on the worker thread while the device imports
        # (and membership-checks) the commitment, which does not depend on the seed
        d = pre[0] if pre is not None else self._seed_finish(self._seed_begin(V.g, V.h), V.u, pkey, w, wp)
        try:
            commitmentTree = V.setCommitment(ByteTreeReader(commitment))
        except EIOException:
            commitmentTree = V.setCommitment(ByteTreeReader(ByteTreeContainer().to_bytes()))
        # ... and so do the replies: everything in the five checks that is a function of the proof alone (two of
        # the three array exponentiations among it) is queued before the seed is known.  A reply that does not
        # parse rejects, as in PoSBasicTW.verify (:1000-1006), once the challenge has been derived.
        ciphPRing = pkey.project(0).getPGroup().getPRing()
        try:
            parsed = V._parseReplies(ciphPRing, ByteTreeReader(reply))
        except EIOException:
            parsed = False
        if parsed:
            V.verifyIndependent()
        prgSeed = d if isinstance(d, bytes) else self.challenger.finish(d)
        V.setBatchVector(prgSeed)
        # the challenge is hashed while the device computes A and F (online: its hash is already running, and
        # stands if the published commitment is exactly the byte tree of what was parsed from it)
        if prc is not None and not (prc[1] is commitment and not getattr(V, "commitmentMalformed", True) and
                                    len(commitment) == commitmentTree.total_bytes() and
                                    bytes(commitment[:5]) == node_header(len(commitmentTree.children))):
            prc[0].abandon()
            prc = None
        if prc is not None:
            cd = prc[0]
        else:
            cd = self.challenger.begin(self.vbitlen)
            ByteTreeContainer(ByteTreeLeaf(prgSeed), commitmentTree).update(cd)
        V.computeAF()
        challengeBytes = self.challenger.finish(cd)
        self.testVector = {"s": prgSeed, "v": _to_positive(challengeBytes)}   # PoS.s / PoS.v of `vmnv -t`
        V.setChallenge(_to_positive(challengeBytes))
        return V.verifyParsed() if parsed else False

    def free(self) -> None:
        self._abandon_prehash()
        for b in (self.P, self.V):
            if b is not None:
                b.free()


class PoSCTW:
    """hvzk/PoSCTW.java:51 (Fiat-Shamir wrapper of PoSCBasicTW) with the bulletin board replaced by
    byte strings: the files PoSCCommitment%02d.bt / PoSCReply%02d.bt of the proof directory (:221-235)."""

    def __init__(self, vbitlen, ebitlen, rbitlen, prg: PRGHeuristic, randomSource, challenger: ChallengerRO):
        self.vbitlen, self.ebitlen, self.rbitlen = vbitlen, ebitlen, rbitlen
        self.prg, self.randomSource, self.challenger = prg, randomSource, challenger

    def _seed(self, g, h, u) -> bytes:
        challengeData = ByteTreeContainer(g.toByteTree(), h.toByteTree(), u.toByteTree())          # :90-92
        return self.challenger.challenge(challengeData, 8 * self.prg.minNoSeedBytes(), self.rbitlen)

    # -- :73-128
    def prove(self, g, h, u, r, pi):
        P = PoSCBasicTW(self.vbitlen, self.ebitlen, self.rbitlen, self.prg, self.randomSource)
        P.setInstance(g, h, u, r, pi)
        prgSeed = self._seed(g, h, u)
        commitment = P.commit(prgSeed)
        challengeData = ByteTreeContainer(ByteTreeLeaf(prgSeed), commitment)
        challengeBytes = self.challenger.challenge(challengeData, self.vbitlen, self.rbitlen)
        reply = P.reply(_to_positive(challengeBytes))
        out = (commitment.to_buffer(), reply.to_buffer())
        P.free()
        return out

    # -- :137-210
    def verify(self, g, h, u, commitment: bytes, reply: bytes) -> bool:
        V = PoSCBasicTW(self.vbitlen, self.ebitlen, self.rbitlen, self.prg, self.randomSource)
        V.setInstance(g, h, u)
        # as in PoSTW.verify: the seed RO(g, h, u) is hashed on the worker thread while the device imports the
        # commitment and the reply and computes what depends on the proof alone
        d = self.challenger.begin(8 * self.prg.minNoSeedBytes())
        ByteTreeContainer(g.toByteTree(), h.toByteTree(), u.toByteTree()).update(d)                # :90-92
        try:
            commitmentTree = V.setCommitment(ByteTreeReader(commitment))
        except EIOException:
            commitmentTree = V.setCommitment(ByteTreeReader(ByteTreeContainer().to_bytes()))
        try:
            parsed = V.parseReply(ByteTreeReader(reply))
        except EIOException:
            parsed = False
        if parsed:
            V.verifyIndependent()
        prgSeed = self.challenger.finish(d)
        V.setBatchVector(prgSeed)
        challengeData = ByteTreeContainer(ByteTreeLeaf(prgSeed), commitmentTree)
        challengeBytes = self.challenger.challenge(challengeData, self.vbitlen, self.rbitlen)
        self.testVector = {"s": prgSeed, "v": _to_positive(challengeBytes)}   # PoSC.s / PoSC.v of `vmnv -t`
        V.setChallenge(_to_positive(challengeBytes))
        verdict = V.verifyParsed() if parsed else False
        V.free()
        return verdict


class CCPoSW:
    """hvzk/CCPoSW.java:53 (Fiat-Shamir wrapper of CCPoSBasicW): CCPoSCommitment%02d.bt /
    CCPoSReply%02d.bt (:274-288).  The verifier side is the plain variant (raisedu == null), which is
    what the stand-alone verifier runs (mixnet/MixNetElGamalVerifyFiatShamirSession.java:757-841)."""

    def __init__(self, vbitlen, ebitlen, rbitlen, prg: PRGHeuristic, randomSource, challenger: ChallengerRO):
        self.vbitlen, self.ebitlen, self.rbitlen = vbitlen, ebitlen, rbitlen
        self.prg, self.randomSource, self.challenger = prg, randomSource, challenger

    def _seed(self, g, h, u, pkey, w, wp) -> bytes:
        challengeData = ByteTreeContainer(g.toByteTree(), h.toByteTree(), u.toByteTree(), pkey.toByteTree(),
                                          w.toByteTree(), wp.toByteTree())                          # :92-98
        return self.challenger.challenge(challengeData, 8 * self.prg.minNoSeedBytes(), self.rbitlen)

    # -- :75-150
    def prove(self, g, h, u, pkey, w, wp, r, pi, s):
        P = CCPoSBasicW(self.vbitlen, self.ebitlen, self.rbitlen, self.prg)
        P.setInstance(g, h, u, pkey, w, wp, r, pi, s)
        # the seed RO(g, h, u, pk, w, w') is hashed on the worker thread (370 MB per 10^5 ciphertexts of width 3)
        # while the device computes the commitment, which does not depend on it
        d = self.challenger.begin(8 * self.prg.minNoSeedBytes())
        ByteTreeContainer(g.toByteTree(), h.toByteTree(), u.toByteTree(), pkey.toByteTree(), w.toByteTree(),
                          wp.toByteTree()).update(d)                                                # :92-98
        P.commitIndependent(self.randomSource)
        prgSeed = self.challenger.finish(d)
        commitment = P.commit(prgSeed, self.randomSource)
        challengeData = ByteTreeContainer(ByteTreeLeaf(prgSeed), commitment)
        challengeBytes = self.challenger.challenge(challengeData, self.vbitlen, self.rbitlen)
        reply = P.reply(_to_positive(challengeBytes))
        out = (commitment.to_buffer(), reply.to_buffer())
        P.free()
        return out

    # -- verifier, online (cf. PoSTW.prehashSeed): the other mix-servers read a party's output from the bulletin board
    # as soon as it is published (mixnet/ShufflerElGamalSession.java:875-890) -- the seed hash starts then, on the
    # premise that the message is well formed (its bytes ARE the byte tree of the parsed array); `verify` checks the
    # premise and hashes again if it does not hold, so verdicts are those of the offline order.
    def prehashSeed(self, g, h, u, pkey, w, output) -> None:
        d = self.challenger.begin(8 * self.prg.minNoSeedBytes())
        d.update(node_header(6))
        for t in (g, h, u, pkey, w):
            t.toByteTree().update(d)
        d.update(output)
        self._preSeed = (d, output)

    def _abandon_prehash(self) -> None:
        pre, self._preSeed = getattr(self, "_preSeed", None), None
        if pre is not None:
            pre[0].abandon()

    # -- :160-260
    def verify(self, g, h, u, pkey, w, wp, commitment: bytes, reply: bytes, outputBytes=None) -> bool:
        """`outputBytes`: the published message `wp` was parsed from (validates a streamed seed hash: it stands only
        if it covered these very bytes and they are the canonical serialisation of `wp`)."""
        V = CCPoSBasicW(self.vbitlen, self.ebitlen, self.rbitlen, self.prg)
        V.setInstance(g, h, u, pkey, w, wp)
        pre, self._preSeed = getattr(self, "_preSeed", None), None
        if pre is not None and not (pre[1] is outputBytes and len(outputBytes) == wp.toByteTree().total_bytes()):
            pre[0].abandon()
            pre = None
        if pre is not None:
            d = pre[0]
        else:
            # as in PoSTW.verify: the seed RO(g, h, u, pk, w, w') is hashed on the worker thread while the device
            # computes what depends on the proof alone (the two multi-exponentiations with k_E)
            d = self.challenger.begin(8 * self.prg.minNoSeedBytes())
            ByteTreeContainer(g.toByteTree(), h.toByteTree(), u.toByteTree(), pkey.toByteTree(), w.toByteTree(),
                              wp.toByteTree()).update(d)                                            # :92-98
        try:
            commitmentTree = V.setCommitment(ByteTreeReader(commitment))
        except EIOException:
            commitmentTree = V.setCommitment(ByteTreeReader(ByteTreeContainer().to_bytes()))
        try:
            parsed = V.parseReply(ByteTreeReader(reply))
        except EIOException:
            parsed = False
        if parsed:
            V.verifyIndependent()
        prgSeed = self.challenger.finish(d)
        V.setBatchVector(prgSeed)
        cd = self.challenger.begin(self.vbitlen)
        ByteTreeContainer(ByteTreeLeaf(prgSeed), commitmentTree).update(cd)
        V.computeAB()
        challengeBytes = self.challenger.finish(cd)
        self.testVector = {"s": prgSeed, "v": _to_positive(challengeBytes)}   # CCPoS.s / CCPoS.v of `vmnv -t`
        V.setChallenge(_to_positive(challengeBytes))
        verdict = V.verifyParsed() if parsed else False
        V.free()
        return verdict
