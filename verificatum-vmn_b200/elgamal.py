"""Decryption factors and their batched proof on the engine: mirror of
`com.verificatum.protocol.elgamal.DistrElGamalSessionBasic` (elgamal/DistrElGamalSessionBasic.java:59)
and of the two array steps of `DistrElGamalSession` (elgamal/DistrElGamalSession.java:377-385, :406).

Array work (one call into the C ABI each):
    decryptionFactors        u.exp(-x*c)              common-exponent variable-base exp   (DistrElGamalSession.java:384-385)
    combineDecryptionFactors pGroup.expProd(bases, integers, bitLength)                    (DistrElGamalSessionBasic.java:502)
    batchInput / batch / batchCombined   expProd with the PRG-derived batching vector      (:524-526, :683-685, :707-709)
The sigma protocol itself is O(k) single-element operations.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

from .arithm import ArithmFormatException, LargeIntegerArray, PFieldElement, expMany
from .eio import ByteTreeBasic, ByteTreeContainer, ByteTreeReader, EIOException
from .hvzk import ProtocolError

# ODD_PRIME_TABLE (elgamal/DistrElGamalSessionBasic.java:199-215): the odd primes up to 1009
ODD_PRIME_TABLE = [n for n in range(3, 1010, 2) if all(n % d for d in range(3, int(n ** 0.5) + 1, 2))]


def primeLog(number: int, prime: int) -> int:
    """:290-302 -- the largest power of `prime` not exceeding `number`."""
    resA = resB = 1
    while resB <= number:
        resA = resB
        resB = resB * prime
    return resA


def prodFactor(pField, k: int) -> PFieldElement:
    """:316-345."""
    maxParties = ODD_PRIME_TABLE[-1]
    if k > maxParties:
        raise ProtocolError("Too many parties! (%d, but at most %d is allowed.)" % (k, maxParties))
    res, prime, i = 1, 2, 0
    while prime <= k:
        res = res * primeLog(k, prime)
        prime = ODD_PRIME_TABLE[i]
        i += 1
    return pField.toElement(res * res)


def modifiedLagrangeCoefficient(pField, prodFactor_: PFieldElement, correct: Sequence[bool], k: int, threshold: int,
                                i: int) -> int:
    """:406-452 -- an *integer* of smallest absolute value (may be negative)."""
    q = pField.group.q
    res = prodFactor_.value
    t = 0
    l = 1
    while t < threshold and l <= k:
        if correct[l]:
            if l != i:
                res = res * l % q
                res = res * pow((l - i) % q, -1, q) % q
            t += 1
        l += 1
    alt = res - q
    return alt if abs(alt) < res else res


def modifiedLagrangeCoefficients(pField, correct: Sequence[bool], k: int, threshold: int) -> List[int]:
    """:362-391."""
    pf = prodFactor(pField, k)
    integers = []
    i = 1
    while len(integers) < threshold and i <= k:
        if correct[i]:
            integers.append(modifiedLagrangeCoefficient(pField, pf, correct, k, threshold, i))
        i += 1
    if len(integers) < threshold:
        raise ProtocolError("Attempting to combine too few decryption factors!")
    return integers


def decryptionFactors(firstComponents, x: PFieldElement, k: int):
    """elgamal/DistrElGamalSession.java:377-385: f = u^{-x * c}, c = prodFactor(k)^{-1} -- ONE exponent
    for all first components: the most expensive per-element operation of the mix-net (VAR(L_q))."""
    pField = x.getPRing()
    inverseFactor = pField.toElement(pow(prodFactor(pField, k).value, -1, pField.group.q))
    return firstComponents.exp(x.neg().mul(inverseFactor))


def combineDecryptionFactors(decryptionFactors_: Dict[int, object], correct: Sequence[bool], k: int, threshold: int):
    """:465-503 -- element-wise prod_j f_j^{lambda_j} with small signed integers."""
    bases = []
    i = 1
    while len(bases) < threshold and i <= k:
        if correct[i]:
            bases.append(decryptionFactors_[i])
        i += 1
    if len(bases) < threshold:
        raise ProtocolError("Attempting to combine too few decryption factors!")
    first = bases[0]
    pGroup = first.getPGroup()
    pField = pGroup.basic()[0].getPRing().getPField()
    integers = modifiedLagrangeCoefficients(pField, correct, k, threshold)
    bitLength = max(abs(v).bit_length() for v in integers)
    if hasattr(first, "comps"):  # product group array: component-wise
        comps = [pGroup.factors[c].expProd([b.comps[c] for b in bases], integers, bitLength)
                 for c in range(len(first.comps))]
        return pGroup.product(*comps)
    return pGroup.expProd(bases, integers, bitLength)


class DistrElGamalSessionBasic:
    """elgamal/DistrElGamalSessionBasic.java:59."""

    def __init__(self, j: int, k: int, threshold: int, ebitlen: int, rbitlen: int, prg):
        self.j, self.k, self.threshold, self.ebitlen, self.rbitlen, self.prg = j, k, threshold, ebitlen, rbitlen, prg
        self.yp: Dict[int, object] = {}
        self.B: Dict[int, object] = {}
        self.Bp: Dict[int, object] = {}
        self.k_x: Dict[int, object] = {}
        self.verdicts = [True] * (k + 1)
        self.e = None

    # :254-283
    def setInstance(self, g, u, y: Dict[int, object], f: Dict[int, object], x: Optional[PFieldElement] = None,
                    combinedy=None, combinedf=None) -> None:
        self.g, self.u, self.y, self.f, self.x = g, u, dict(y), dict(f), x
        self.combinedy, self.combinedf = combinedy, combinedf
        self.pField = g.getPGroup().getPRing().getPField()
        self.inverseFactor = self.pField.toElement(pow(prodFactor(self.pField, self.k).value, -1, self.pField.group.q))

    # :513-518
    def setBatchVector(self, prgSeed: bytes) -> None:
        self.prg.setSeed(prgSeed)
        lia = LargeIntegerArray.random(self.u.size(), self.ebitlen, self.prg, self.pField)
        self.e = self.pField.unsafeToElementArray(lia)

    # :524-526
    def batchInput(self) -> None:
        self.A = self.u.expProd(self.e)

    # :534-540
    def commit(self, randomSource) -> ByteTreeBasic:
        self.r = self.g.getPGroup().getPRing().randomElement(randomSource, self.rbitlen)
        self.yp[self.j] = self.g.exp(self.r)
        self.Bp[self.j] = self.A.exp(self.r)
        return ByteTreeContainer(self.yp[self.j].toByteTree(), self.Bp[self.j].toByteTree())

    # :549-565
    def setCommitment(self, l: int, commitmentReader) -> None:
        """`commitmentReader`: a ByteTreeReader, or the raw bytes of the file -- the reader's constructor parses
        the first header, so a truncated or empty file must fail INSIDE the handler below (verdicts[l] = false,
        trivial values substituted, the combined check still runs: :553-566)."""
        try:
            if not isinstance(commitmentReader, ByteTreeReader):
                commitmentReader = ByteTreeReader(commitmentReader)
            self.yp[l] = self.g.getPGroup().toElement(commitmentReader.getNextChild())
            self.Bp[l] = self.A.getPGroup().toElement(commitmentReader.getNextChild())
        except (EIOException, ArithmFormatException):
            self.verdicts[l] = False
        if not self.verdicts[l]:
            self.yp[l] = self.g.getPGroup().getONE()
            self.Bp[l] = self.A.getPGroup().getONE()

    # :573-588
    def getCommitment(self, l: Optional[int] = None) -> ByteTreeBasic:
        if l is None:
            return ByteTreeContainer(*[self.getCommitment(i + 1) for i in range(self.k)])
        return ByteTreeContainer(self.yp[l].toByteTree(), self.Bp[l].toByteTree())

    # :595-598
    def reply(self, v: int) -> ByteTreeBasic:
        self.k_x[self.j] = self.x.neg().mul(self.inverseFactor).mul(self.pField.toElement(v)).add(self.r)
        return self.k_x[self.j].toByteTree()

    # :606-614
    def setReply(self, l: int, replyReader) -> None:
        pRing = self.g.getPGroup().getPRing()
        try:
            if not isinstance(replyReader, ByteTreeReader):
                replyReader = ByteTreeReader(replyReader)
            self.k_x[l] = pRing.toElement(replyReader)
        except (EIOException, ArithmFormatException):
            self.k_x[l] = pRing.getZERO()
            self.verdicts[l] = False

    def getReply(self, l: int) -> ByteTreeBasic:
        return self.k_x[l].toByteTree()

    def getVerdict(self, l: int) -> bool:
        return self.verdicts[l]

    # :642-678
    def combine(self, correct: Sequence[bool]) -> None:
        integers = modifiedLagrangeCoefficients(self.pField, correct, self.k, self.threshold)
        exponents = [self.pField.toElement(-v).neg() if v < 0 else self.pField.toElement(v) for v in integers]
        self.combinedyp = self.yp[1].getPGroup().getONE()
        self.combinedBp = self.Bp[1].getPGroup().getONE()
        self.combinedk_x = self.k_x[1].getPRing().getZERO()
        t, l, used = 0, 1, []
        while t < self.threshold and l <= self.k:
            if correct[l]:
                used.append((l, t))
                t += 1
            l += 1
        # the 2|S| single-element exponentiations are independent: one small array call (arithm.expMany)
        powers = expMany([self.yp[l] for l, _ in used] + [self.Bp[l] for l, _ in used],
                         [exponents[t] for _, t in used] * 2)
        for j, (l, t) in enumerate(used):
            self.combinedyp = self.combinedyp.mul(powers[j])
            self.combinedBp = self.combinedBp.mul(powers[len(used) + j])
            self.combinedk_x = self.combinedk_x.add(self.k_x[l].mul(exponents[t]))

    # :683-685
    def batchCombined(self) -> None:
        self.combinedB = self.combinedf.expProd(self.e)

    # :693-700
    def verifyCombined(self, v: int) -> bool:
        pfev = self.pField.toElement(v)
        yv, Bv, Ak = expMany([self.combinedy.inv(), self.combinedB, self.A], [pfev, pfev, self.combinedk_x])
        return (yv.mul(self.combinedyp).equals(self.g.exp(self.combinedk_x))
                and Bv.mul(self.combinedBp).equals(Ak))

    # :707-709
    def batch(self, l: int) -> None:
        self.B[l] = self.f[l].expProd(self.e)

    # :718-727
    def verify(self, l: int, v: int) -> bool:
        if not self.verdicts[l]:
            return False
        pfev = self.pField.toElement(v)
        return (self.y[l].inv().exp(self.inverseFactor.mul(pfev)).mul(self.yp[l]).equals(self.g.exp(self.k_x[l]))
                and self.B[l].exp(pfev).mul(self.Bp[l]).equals(self.A.exp(self.k_x[l])))

    def free(self) -> None:
        if self.e is not None:
            self.e.free()
            self.e = None
