"""Re-encryption shuffle on the engine: mirror of the arithmetic in
`mixnet/ShufflerElGamalSession.java` (shuffle:362-433, performShuffling:250-350) together with
the helpers that define the ciphertext layout (`elgamal/ProtocolElGamal.java:753-800`), the
independent generators (`distr/IndependentGeneratorsRO.java:110-130`), the global prefix
(`elgamal/ProtocolElGamal.java:659-683`) and the demo ciphertext generator
(`elgamal/ProtocolElGamalInterfaceRaw.java:99-130`).

The bulletin board, XML configuration, state files and the k-party choreography are out of
scope: one `ShufflerSession` object plays either the proving mix-server
(`shuffle`) or a verifying one (`verify`) on in-memory byte trees.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from .arithm import ModPGroup, Permutation, PPGroup, PPGroupElement
from .crypto import HashfunctionHeuristic, PRGHeuristic, RandomOracle
from .arithm import ArithmFormatException
from .eio import ByteTreeContainer, ByteTreeLeaf, ByteTreeReader, EIOException, booleanArrayToByteTree, int32_leaf
from .hvzk import CCPoSW, ChallengerRO, PoSCTW, PoSTW


# ---------------------------------------------------------------- elgamal/ProtocolElGamal.java:753-800
def getPlainPGroup(pGroup, width: int):
    return pGroup if width == 1 else PPGroup(pGroup, width)


def getCiphPGroup(pGroup, width: int) -> PPGroup:
    if width == 1:
        return PPGroup(pGroup, 2)
    return PPGroup(PPGroup(pGroup, width), 2)


def getWidePublicKey(fullPublicKey: PPGroupElement, width: int) -> PPGroupElement:
    if width == 1:
        return fullPublicKey
    g = fullPublicKey.project(0)
    y = fullPublicKey.project(1)
    ciphPGroup = getCiphPGroup(g.getPGroup(), width)
    plainPGroup = ciphPGroup.project(0)
    return ciphPGroup.product(plainPGroup.product(g), plainPGroup.product(y))


# ---------------------------------------------------------------- elgamal/ProtocolElGamal.java:659-683
def globalPrefix(roHashfunction: HashfunctionHeuristic, version: str, rosid: str, rbitlen: int, vbitlenro: int,
                 ebitlenro: int, prgString: str, pGroupString: str, roHashfunctionString: str) -> bytes:
    bt = ByteTreeContainer(ByteTreeLeaf(version.encode()), ByteTreeLeaf(rosid.encode()), int32_leaf(rbitlen),
                           int32_leaf(vbitlenro), int32_leaf(ebitlenro), ByteTreeLeaf(prgString.encode()),
                           ByteTreeLeaf(pGroupString.encode()), ByteTreeLeaf(roHashfunctionString.encode()))
    return roHashfunction.hash(bt.to_bytes())


# ---------------------------------------------------------------- distr/IndependentGeneratorsRO.java:110-130
class IndependentGeneratorsRO:
    def __init__(self, sid: str, roHashfunction: HashfunctionHeuristic, globalPrefix: bytes, rbitlen: int):
        self.sid, self.roHashfunction, self.globalPrefix, self.rbitlen = sid, roHashfunction, globalPrefix, rbitlen

    def generate(self, pGroup: ModPGroup, numberOfGenerators: int):
        prg = PRGHeuristic(self.roHashfunction)
        ro = RandomOracle(self.roHashfunction, 8 * prg.minNoSeedBytes())
        d = ro.getDigest()
        d.update(self.globalPrefix)
        d.update(ByteTreeLeaf(self.sid.encode()).to_bytes())
        prg.setSeed(d.digest())
        return pGroup.randomElementArray(numberOfGenerators, prg, self.rbitlen)


# ---------------------------------------------------------------- elgamal/ProtocolElGamalInterfaceRaw.java:99-130
def demoCiphertexts(fullPublicKey: PPGroupElement, noCiphs: int, randomSource):
    basicPublicKey = fullPublicKey.project(0)
    publicKey = fullPublicKey.project(1)
    pRing = publicKey.getPGroup().getPRing()
    m = publicKey.getPGroup().randomElementArray(noCiphs, randomSource, 10)
    r = pRing.randomElementArray(noCiphs, randomSource, 20)
    u = basicPublicKey.exp(r)
    t = publicKey.exp(r)
    r.free()
    v = t.mul(m)
    t.free()
    m.free()
    return fullPublicKey.getPGroup().product(u, v)


@dataclass
class SessionParams:
    """The protocol parameters the hot path needs (elgamal/ProtocolElGamalGen.java:81-213)."""
    vbitlenro: int = 256
    ebitlenro: int = 256
    rbitlen: int = 100
    rohash: str = "SHA-256"
    prghash: str = "SHA-256"
    version: str = "3.1.0"
    sid: str = "vmx"              # session identifier of the protocol info file
    auxsid: str = "default"       # auxiliary session identifier of this execution (the `auxsid` file of a proof)
    pGroupString: str = ""

    @property
    def rosid(self) -> str:
        """mixnet/MixNetElGamalVerifyFiatShamirSession.java:160: sid + "." + auxsid goes into the global prefix."""
        return self.sid + "." + self.auxsid


def validateSid(sid: str) -> bool:
    """[VCR-mem] Protocol.validateSid: a non-empty string of letters, digits, underscores and spaces."""
    import re
    return re.fullmatch(r"[A-Za-z0-9_ ]{1,1024}", sid) is not None


@dataclass
class ShuffleProof:
    """What one mix-server publishes for one shuffle: the files Ciphertexts%02d.bt,
    PermutationCommitment%02d.bt, PoSCommitment%02d.bt, PoSReply%02d.bt of the proof directory
    (mixnet/MixNetElGamalSession.java:381-446, hvzk/PoSTW.java:281-307).  The fields are immutable bytes-like
    values: `bytes`, or `eio.HostBytes` over a page-locked buffer when the message is large (ByteTreeBasic.to_buffer)."""
    output: bytes
    permutationCommitment: bytes
    commitment: bytes
    reply: bytes


class ShufflerSession:
    """The arithmetic of ShufflerElGamalSession for one party."""

    def __init__(self, pGroup: ModPGroup, publicKey: PPGroupElement, params: SessionParams, randomSource,
                 sid: str = "1"):
        self.pGroup = pGroup
        self.publicKey = publicKey
        self.params = params
        self.randomSource = randomSource
        self.roHashfunction = HashfunctionHeuristic(params.rohash)
        self.prg = PRGHeuristic(HashfunctionHeuristic(params.prghash))
        self.globalPrefix = globalPrefix(self.roHashfunction, params.version, params.rosid, params.rbitlen,
                                         params.vbitlenro, params.ebitlenro, "PRGHeuristic(%s)" % params.prghash,
                                         params.pGroupString, "HashfunctionHeuristic(%s)" % params.rohash)
        self.challenger = ChallengerRO(self.roHashfunction, self.globalPrefix, comm=getattr(pGroup, "comm", None))
        self.sid = sid
        self.generators = None

    def _pos(self) -> PoSTW:
        p = self.params
        return PoSTW(p.vbitlenro, p.ebitlenro, p.rbitlen, self.prg, self.randomSource, self.challenger)

    # ShufflerElGamalSession.java:384
    def deriveGenerators(self, size: int):
        igs = IndependentGeneratorsRO("generators", self.roHashfunction, self.globalPrefix, self.params.rbitlen)
        return igs.generate(self.pGroup, size)

    # ShufflerElGamalSession.java:362-433 + :250-300 for l == j
    def shuffle(self, width: int, ciphertexts, generators=None, keep_output: bool = False, publish=None):
        """Re-encrypt, permute and prove.  Returns (ShuffleProof, output array or None).  `publish(name, message)`
        is called as each of the four messages ("output", "permutationCommitment", "commitment", "reply") goes
        to the bulletin board (an OnlineVerification's `publish`, a file writer)."""
        ciphPPGroup = ciphertexts.getPGroup()
        exponentsPRing = ciphPPGroup.project(0).getPRing()
        widePublicKey = getWidePublicKey(self.publicKey, width)
        size = ciphertexts.size()
        own_generators = generators is None
        if own_generators:
            generators = self.deriveGenerators(size)
        rbitlen = self.params.rbitlen
        P = self._pos()
        P.beginSeed(generators.getPGroup().getg(), generators)   # Fiat-Shamir hashing of (g, h) runs beside the GPU
        # The random source is consumed in the reference's order (exponents, permutation, then the proof's r,
        # alpha, epsilon), but the device computes u BEFORE the re-encryption factors: u is the third input of
        # the Fiat-Shamir seed, and its hashing -- then that of pk and w -- runs beside the re-encryption.
        reencExponents = exponentsPRing.randomElementArray(size, self.randomSource, rbitlen)       # :400-403
        permutation = Permutation.random(size, self.randomSource, rbitlen, self.pGroup.basic()[0])                       # :408-409
        P.precompute(generators.getPGroup().getg(), generators, permutation)                       # :414
        reencFactors = widePublicKey.exp(reencExponents)                                           # :407
        P.continueSeed(widePublicKey, ciphertexts)
        reenc = ciphertexts.mul(reencFactors)                                                      # :273
        reencFactors.free()
        inverse = permutation.inv()
        output = reenc.permute(inverse)                                                            # :278
        reenc.free()
        inverse.free()
        output_bytes = output.toByteTree().to_buffer()                                              # :284
        if publish is not None:
            publish("output", output_bytes)
        pc, commitment, reply = P.prove(widePublicKey, ciphertexts, output, reencExponents, publish)   # :289
        reencExponents.free()
        if own_generators:
            generators.free()
        permutation.free()
        proof = ShuffleProof(output_bytes, pc, commitment, reply)
        if keep_output:
            return proof, output
        output.free()
        return proof, None

    # ShufflerElGamalSession.java:195-210 (readOutput) + :301-330 (verify branch)
    def beginVerify(self, width: int, ciphertexts, generators=None) -> "OnlineVerification":
        """Verification of a shuffle WHILE it is being proved: the returned object's `publish` is handed to the
        prover (`shuffle(..., publish=...)`), `finish(proof)` gives the verdict.  This is how the reference's
        mix-servers verify each other -- they wait on the bulletin board message by message (hvzk/PoSTW.java:195-245,
        mixnet/ShufflerElGamalSession.java:301-330) -- and it takes the verifier's Fiat-Shamir hashing (3.1 KB per
        ciphertext at 3072 bits, one SHA-256 stream) off the critical path: it runs beside the prover's."""
        return OnlineVerification(self, width, ciphertexts, generators)

    def verify(self, width: int, ciphertexts, proof: ShuffleProof, generators=None, output=None, _pos=None):
        """Returns (verdict, output array) -- on failure the output is a copy of the input
        ("Replacing output with input", :321-327).  `output`: proof.output already parsed (vmnv reads it itself)."""
        ciphPPGroup = ciphertexts.getPGroup()
        widePublicKey = getWidePublicKey(self.publicKey, width)
        size = ciphertexts.size()
        own_generators = generators is None
        if own_generators:
            generators = self.deriveGenerators(size)
        outputBytes = None
        try:
            if output is None:
                output = ciphPPGroup.toElementArray(size, ByteTreeReader(proof.output))
                outputBytes = proof.output
        except Exception:
            if own_generators:
                generators.free()
            if _pos is not None:
                _pos.free()
            return False, ciphertexts.copyOfRange(0, size)
        V = _pos
        if V is None:
            V = self._pos()
            V.precompute(generators.getPGroup().getg(), generators)
        verdict = V.verify(widePublicKey, ciphertexts, output, proof.permutationCommitment, proof.commitment,
                           proof.reply, outputBytes=outputBytes)
        self.last_u_parsed = V.u_parsed
        self.last_test_vector = V.testVector
        V.free()
        if own_generators:
            generators.free()
        if not verdict:
            output.free()
            output = ciphertexts.copyOfRange(0, size)
        return verdict, output


class OnlineVerification:
    """One shuffle being verified as its messages are published (ShufflerSession.beginVerify)."""

    def __init__(self, session: ShufflerSession, width: int, ciphertexts, generators=None):
        self.session, self.width, self.ciphertexts = session, width, ciphertexts
        self.own_generators = generators is None
        self.generators = session.deriveGenerators(ciphertexts.size()) if generators is None else generators
        self.widePublicKey = getWidePublicKey(session.publicKey, width)
        self.V = session._pos()
        self.V.precompute(self.generators.getPGroup().getg(), self.generators)
        self.messages = {}

    def publish(self, name: str, message) -> None:
        """A message appeared on the board: start the hashing it unlocks (no verdict depends on this being
        called -- `finish` redoes whatever was not, or not validly, anticipated)."""
        self.messages[name] = message
        m = self.messages
        if name in ("output", "permutationCommitment") and "output" in m and "permutationCommitment" in m:
            self.V.prehashSeed(self.widePublicKey, self.ciphertexts, m["permutationCommitment"], m["output"])
        elif name == "commitment":
            self.V.prehashChallenge(message)

    def finish(self, proof: ShuffleProof):
        """(verdict, output array), as ShufflerSession.verify."""
        verdict, out = self.session.verify(self.width, self.ciphertexts, proof, generators=self.generators, _pos=self.V)
        if self.own_generators:
            self.generators.free()
        return verdict, out


# ---------------------------------------------------------------- mixnet/PermutationCommitment.java
class PermutationCommitment:
    """Pre-computed commitment u = (h * g^t) permuted of one party (mixnet/PermutationCommitment.java:56).

    Array work: g.exp(exponents) fixed-base (:200), mul (:201), permute (:215), the PoSC proof
    (:286-291 / :309-313), extract(keepList) when the commitment is shrunk to the actual number of
    ciphertexts (:462-468).  State files and the bulletin board are replaced by attributes / bytes."""

    def __init__(self, session: "ShufflerSession", generators):
        self.session = session
        self.generators = generators
        self.pGroup = generators.getPGroup()
        self.exponents = self.identityCommitment = self.permutation = self.commitment = None
        self.raisedCommitment = None

    def _posc(self) -> PoSCTW:
        p, s = self.session.params, self.session
        return PoSCTW(p.vbitlenro, p.ebitlenro, p.rbitlen, s.prg, s.randomSource, s.challenger)

    # -- :148-219 (the branch that generates fresh values)
    def precompute(self) -> None:
        s = self.session
        size = self.generators.size()
        self.exponents = self.pGroup.getPRing().randomElementArray(size, s.randomSource, s.params.rbitlen)
        tmp = self.pGroup.getg().exp(self.exponents)
        self.identityCommitment = self.generators.mul(tmp)
        tmp.free()
        self.permutation = Permutation.random(size, s.randomSource, s.params.rbitlen, self.pGroup.basic()[0])
        self.commitment = self.identityCommitment.permute(self.permutation)

    # -- :251-292 for l == j: publish the commitment and prove knowledge of (exponents, permutation)
    def prove(self):
        """Returns (PermutationCommitment%02d.bt, PoSCCommitment%02d.bt, PoSCReply%02d.bt)."""
        c, r = self._posc().prove(self.pGroup.getg(), self.generators, self.commitment, self.exponents,
                                  self.permutation)
        return self.commitment.toByteTree().to_buffer(), c, r

    # -- :293-346 for l != j
    def verify(self, commitment: bytes, poscCommitment: bytes, poscReply: bytes, raisedExponent=None) -> bool:
        """Reads another party's commitment and verifies its PoSC; a malformed commitment or a
        rejected proof replaces it with the generators ("trivial commitment of identity permutation")."""
        size = self.generators.size()
        trivial = False
        try:
            self.commitment = self.pGroup.toElementArray(size, ByteTreeReader(commitment))
        except (ArithmFormatException, EIOException):
            trivial = True
        if not trivial:
            trivial = not self._posc().verify(self.pGroup.getg(), self.generators, self.commitment, poscCommitment,
                                              poscReply)
            if trivial:
                self.commitment.free()
        if trivial:
            self.commitment = self.generators.copyOfRange(0, size)
        if raisedExponent is not None:
            self.raisedCommitment = self.commitment.exp(raisedExponent)                             # :357
        return not trivial

    # -- :390-471
    def shrink(self, noCiphertexts: int, keepList: Optional[bytes] = None) -> bytes:
        """l == j (keepList is None): derive and publish the keep list, shrink exponents and permutation.
        l != j: read the published keep list (trivial list on any defect).  Returns KeepList%02d.bt."""
        import numpy as np
        size = self.commitment.size()
        if keepList is None:
            keep = np.zeros(size, dtype=bool)
            keep[self.permutation.table[:noCiphertexts]] = True
            old = self.exponents
            self.exponents = old.copyOfRange(0, noCiphertexts)
            old.free()
            self.permutation = self.permutation.shrink(noCiphertexts)
        else:
            trivial = False
            try:
                keep = ByteTreeReader(keepList).readBooleans(size)
            except EIOException:
                trivial = True
            if not trivial and int(keep.sum()) != noCiphertexts:
                trivial = True
            if trivial:
                keep = np.zeros(size, dtype=bool)
                keep[:noCiphertexts] = True
        old = self.commitment
        self.commitment = old.extract(keep)
        old.free()
        if self.raisedCommitment is not None:
            old = self.raisedCommitment
            self.raisedCommitment = old.extract(keep)
            old.free()
        return booleanArrayToByteTree(keep).to_bytes()

    def free(self) -> None:
        for a in (self.exponents, self.identityCommitment, self.commitment, self.raisedCommitment):
            if a is not None:
                a.free()
        self.exponents = self.identityCommitment = self.commitment = self.raisedCommitment = None


@dataclass
class CommittedShuffleProof:
    """What a mix-server publishes for one commitment-consistent shuffle: Ciphertexts%02d.bt,
    CCPoSCommitment%02d.bt, CCPoSReply%02d.bt (hvzk/CCPoSW.java:274-288)."""
    output: bytes
    commitment: bytes
    reply: bytes


class CommittedShuffler:
    """Pre-computation and commitment-consistent shuffling of one party: the arithmetic of
    ShufflerElGamalSession.precomp (:534-672), shrink (:673-760) and committedShuffle (:771-960)."""

    def __init__(self, session: "ShufflerSession", width: int, maxciph: int):
        self.session = session
        self.width = width
        self.maxciph = maxciph
        self.generators = session.deriveGenerators(maxciph)                                          # :568
        self.widePublicKey = getWidePublicKey(session.publicKey, width)
        self.permutationCommitment = PermutationCommitment(session, self.generators)
        self.reencExponents = self.reencFactors = None

    def _ccpos(self) -> CCPoSW:
        p, s = self.session.params, self.session
        return CCPoSW(p.vbitlenro, p.ebitlenro, p.rbitlen, s.prg, s.randomSource, s.challenger)

    # -- :598-658 for an active party
    def precomp(self):
        """Commit to a permutation and pre-compute the re-encryption factors for maxciph ciphertexts.
        Returns the three published byte strings of PermutationCommitment.prove()."""
        s = self.session
        self.permutationCommitment.precompute()
        published = self.permutationCommitment.prove()
        exponentsPRing = getCiphPGroup(s.pGroup, self.width).project(0).getPRing()
        self.reencExponents = exponentsPRing.randomElementArray(self.maxciph, s.randomSource, s.params.rbitlen)  # :647-651
        self.reencFactors = self.widePublicKey.exp(self.reencExponents)                                          # :658
        return published

    # -- :673-760
    def shrink(self, noCiphertexts: int) -> bytes:
        old = self.generators
        self.generators = old.copyOfRange(0, noCiphertexts)
        self.permutationCommitment.generators = self.generators
        old.free()
        if self.reencExponents is not None:
            old = self.reencExponents
            self.reencExponents = _copy_of_range(old, 0, noCiphertexts)
            old.free()
            old = self.reencFactors
            self.reencFactors = old.copyOfRange(0, noCiphertexts)
            old.free()
        return self.permutationCommitment.shrink(noCiphertexts)

    # -- :771-822
    def shuffle(self, ciphertexts, keep_output: bool = False, publish=None):
        """`publish(name, message)` is called as the output goes to the bulletin board (:795), before it is proved
        (an OnlineCommittedVerification's `publish`)."""
        pc = self.permutationCommitment
        reenc = ciphertexts.mul(self.reencFactors)                                                   # :789
        inverse = pc.permutation.inv()
        output = reenc.permute(inverse)                                                              # :792
        reenc.free()
        output_bytes = output.toByteTree().to_buffer()
        if publish is not None:
            publish("output", output_bytes)
        c, r = self._ccpos().prove(self.generators.getPGroup().getg(), self.generators, pc.commitment,
                                   self.widePublicKey, ciphertexts, output, pc.exponents, pc.permutation,
                                   self.reencExponents)                                              # :808-819
        proof = CommittedShuffleProof(output_bytes, c, r)
        if keep_output:
            return proof, output
        output.free()
        return proof, None


def _copy_of_range(arr, a: int, b: int):
    if hasattr(arr, "comps"):
        from .arithm import PPRingElementArray
        return PPRingElementArray(arr.ring, [c.copyOfRange(a, b) for c in arr.comps])
    return arr.copyOfRange(a, b)


class OnlineCommittedVerification:
    """One commitment-consistent shuffle being verified as its messages are published: the verifier's seed hash
    (h, u, pk, w, w' -- 370 MB per 10^5 ciphertexts of width 3) starts when the output appears on the board and
    runs beside the prover's (mixnet/ShufflerElGamalSession.java:875-960: the other parties read the output as soon
    as it is published).  `publish` goes to `CommittedShuffler.shuffle(..., publish=...)`; `finish(proof)` gives
    (verdict, output array) exactly as `verifyCommittedShuffle` would."""

    def __init__(self, session: "ShufflerSession", width: int, generators, commitment, ciphertexts):
        self.session, self.width, self.generators = session, width, generators
        self.commitment, self.ciphertexts = commitment, ciphertexts
        p = session.params
        self.V = CCPoSW(p.vbitlenro, p.ebitlenro, p.rbitlen, session.prg, session.randomSource, session.challenger)

    def publish(self, name: str, message) -> None:
        if name == "output":
            self.V.prehashSeed(self.generators.getPGroup().getg(), self.generators, self.commitment,
                               getWidePublicKey(self.session.publicKey, self.width), self.ciphertexts, message)

    def finish(self, proof: CommittedShuffleProof):
        return verifyCommittedShuffle(self.session, self.width, self.generators, self.commitment, self.ciphertexts,
                                      proof, _ccpos=self.V)


def verifyCommittedShuffle(session: "ShufflerSession", width: int, generators, commitment, ciphertexts,
                           proof: CommittedShuffleProof, _ccpos=None):
    """ShufflerElGamalSession.committedShuffleVerify (:875-960) without the raised-commitment
    optimisation: read the output, verify the CCPoS against the (shrunk) permutation commitment; on
    failure the output is replaced by the input.  Returns (verdict, output)."""
    ciphPPGroup = ciphertexts.getPGroup()
    size = ciphertexts.size()
    widePublicKey = getWidePublicKey(session.publicKey, width)
    V = _ccpos
    try:
        output = ciphPPGroup.toElementArray(size, ByteTreeReader(proof.output))
    except Exception:
        if V is not None:
            V._abandon_prehash()
        return False, ciphertexts.copyOfRange(0, size)
    p = session.params
    if V is None:
        V = CCPoSW(p.vbitlenro, p.ebitlenro, p.rbitlen, session.prg, session.randomSource, session.challenger)
    verdict = V.verify(generators.getPGroup().getg(), generators, commitment, widePublicKey, ciphertexts, output,
                       proof.commitment, proof.reply, outputBytes=proof.output)
    if not verdict:
        output.free()
        output = ciphertexts.copyOfRange(0, size)
    return verdict, output
