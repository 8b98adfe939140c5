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
from .eio import ByteTreeContainer, ByteTreeLeaf, ByteTreeReader, int32_leaf
from .hvzk import ChallengerRO, PoSTW


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
    rosid: str = "vmx.session"
    pGroupString: str = ""


@dataclass
class ShuffleProof:
    """What one mix-server publishes for one shuffle: the files Ciphertexts%02d.bt,
    PermutationCommitment%02d.bt, PoSCommitment%02d.bt, PoSReply%02d.bt of the proof directory
    (mixnet/MixNetElGamalSession.java:381-446, hvzk/PoSTW.java:281-307)."""
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
        self.challenger = ChallengerRO(self.roHashfunction, self.globalPrefix)
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
    def shuffle(self, width: int, ciphertexts, generators=None, keep_output: bool = False):
        """Re-encrypt, permute and prove.  Returns (ShuffleProof, output array or None)."""
        ciphPPGroup = ciphertexts.getPGroup()
        exponentsPRing = ciphPPGroup.project(0).getPRing()
        widePublicKey = getWidePublicKey(self.publicKey, width)
        size = ciphertexts.size()
        own_generators = generators is None
        if own_generators:
            generators = self.deriveGenerators(size)
        rbitlen = self.params.rbitlen
        reencExponents = exponentsPRing.randomElementArray(size, self.randomSource, rbitlen)       # :400-403
        reencFactors = widePublicKey.exp(reencExponents)                                           # :407
        permutation = Permutation.random(size, self.randomSource, rbitlen)                         # :408-409
        P = self._pos()
        P.precompute(generators.getPGroup().getg(), generators, permutation)                       # :414
        reenc = ciphertexts.mul(reencFactors)                                                      # :273
        reencFactors.free()
        inverse = permutation.inv()
        output = reenc.permute(inverse)                                                            # :278
        reenc.free()
        inverse.free()
        output_bytes = output.toByteTree().to_bytes()                                              # :284
        pc, commitment, reply = P.prove(widePublicKey, ciphertexts, output, reencExponents)        # :289
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
    def verify(self, width: int, ciphertexts, proof: ShuffleProof, generators=None):
        """Returns (verdict, output array) -- on failure the output is a copy of the input
        ("Replacing output with input", :321-327)."""
        ciphPPGroup = ciphertexts.getPGroup()
        widePublicKey = getWidePublicKey(self.publicKey, width)
        size = ciphertexts.size()
        own_generators = generators is None
        if own_generators:
            generators = self.deriveGenerators(size)
        try:
            output = ciphPPGroup.toElementArray(size, ByteTreeReader(proof.output))
        except Exception:
            if own_generators:
                generators.free()
            return False, ciphertexts.copyOfRange(0, size)
        V = self._pos()
        V.precompute(generators.getPGroup().getg(), generators)
        verdict = V.verify(widePublicKey, ciphertexts, output, proof.permutationCommitment, proof.commitment,
                           proof.reply)
        V.free()
        if own_generators:
            generators.free()
        if not verdict:
            output.free()
            output = ciphertexts.copyOfRange(0, size)
        return verdict, output
