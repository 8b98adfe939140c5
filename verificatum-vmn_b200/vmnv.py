"""A whole mix on the engine, and its universal verification: the array work of

    mixnet/MixNetElGamalSession (shuffle by the first `activeThreshold` parties, then decryption),
    elgamal/DistrElGamalSession.java:361-545 (decryption factors, their exchange, the batched proof),
    mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668 (what `vmnv` does with a proof directory:
    2 x verifyPoS + verification of the decryption, BASELINE.json config 3)

in the order and with the Fiat-Shamir inputs of the reference, on in-memory byte trees named like the
files of the proof directory (mixnet/MixNetElGamalSession.java:381-446, mixnet/ShufflerElGamalSession.java:1077-1101,
hvzk/PoSTW.java:281-307, elgamal/DistrElGamalSession.java:553-601, elgamal/DistrElGamal.java:245-255).

Out of scope: the bulletin board, the distributed key generation (replaced by a Shamir sharing
in the exponent dealt from one random source -- the verifier only ever sees its public outcome,
PolynomialInExponent.bt and FullPublicKey.bt), info files, the command line tools.  The k parties run in one
process, as the reference's own demo does (Demo.java:282-290).

[VCR-mem] PolynomialInExponent.toByteTree() is taken to be the node of its degree + 1 coefficient elements.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional, Sequence

from . import elgamal as eg
from .arithm import ArithmFormatException, PFieldElement
from .crypto import PRGHeuristic
from .eio import ByteTreeContainer, ByteTreeLeaf, ByteTreeReader, EIOException, booleanArrayToByteTree
from .hvzk import _to_positive, node_header
from .mixnet import SessionParams, ShuffleProof, ShufflerSession, getCiphPGroup, validateSid


def _parse_int(raw: bytes) -> int:
    """Integer.parseInt of a file's content: an optional sign and decimal digits, nothing else."""
    import re
    text = raw.decode("ascii", errors="replace")
    if re.fullmatch(r"[+-]?[0-9]{1,10}", text) is None:
        raise ValueError("not an integer: %r" % text[:20])
    return int(text)


class VerificationError(RuntimeError):
    """`v.failStop(...)` of the reference's verifier: the proof directory is unusable."""


# ---------------------------------------------------------------- proof directory
class ProofDirectory(dict):
    """relative file name -> bytes (the nizkp directory of vmn / vmnv)."""

    @staticmethod
    def Lfile(l: int) -> str:
        return "proofs/Ciphertexts%02d.bt" % l

    @staticmethod
    def PCfile(l: int) -> str:
        return "proofs/PermutationCommitment%02d.bt" % l

    @staticmethod
    def PoSCfile(l: int) -> str:
        return "proofs/PoSCommitment%02d.bt" % l

    @staticmethod
    def PoSRfile(l: int) -> str:
        return "proofs/PoSReply%02d.bt" % l

    @staticmethod
    def DFfile(l: int) -> str:
        return "proofs/DecryptionFactors%02d.bt" % l

    @staticmethod
    def DFCfile(l: int) -> str:
        return "proofs/DecrFactCommitment%02d.bt" % l

    @staticmethod
    def DFRfile(l: int) -> str:
        return "proofs/DecrFactReply%02d.bt" % l

    def write(self, root: str) -> None:
        for name, data in self.items():
            path = os.path.join(root, name)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "wb") as f:
                f.write(data)

    @staticmethod
    def read(root: str) -> "ProofDirectory":
        d = ProofDirectory()
        for base, _, files in os.walk(root):
            for fn in files:
                path = os.path.join(base, fn)
                with open(path, "rb") as f:
                    d[os.path.relpath(path, root).replace(os.sep, "/")] = f.read()
        return d


def _evaluate_in_exponent(coeffs, l: int):
    """PolynomialInExponent.evaluate(l) = prod_i c_i^(l^i) (single elements, through the engine)."""
    pField = coeffs[0].getPGroup().getPRing()
    acc = coeffs[0]
    power = 1
    for c in coeffs[1:]:
        power = power * l
        acc = acc.mul(c.exp(pField.toElement(power)))
    return acc


# ---------------------------------------------------------------- the mix (k parties in one process)
class MixNetElGamal:
    """Keys + mixing + decryption of one list of ciphertexts; everything published goes to `self.nizkp`."""

    def __init__(self, pGroup, params: SessionParams, k: int, threshold: int, randomSource, width: int = 1,
                 auxsid: Optional[str] = None):
        if auxsid is not None:
            params = dataclasses.replace(params, auxsid=auxsid)
        if width < 1:
            raise ValueError("width must be positive")
        self.pGroup, self.params, self.k, self.threshold, self.width = pGroup, params, k, threshold, width
        self.randomSource = randomSource
        pField = pGroup.getPRing()
        # Shamir sharing of the joint secret key in the exponent (stands in for the DKG of the control plane)
        self.poly = [pField.randomElement(randomSource, params.rbitlen) for _ in range(threshold)]
        self.secretKeys = {}
        for l in range(1, k + 1):
            acc, power = pField.getZERO(), 1
            for a in self.poly:
                acc = acc.add(a.mul(pField.toElement(power)))
                power *= l
            self.secretKeys[l] = acc
        g = pGroup.getg()
        self.polynomialInExponent = [g.exp(a) for a in self.poly]
        self.publicKeys = {l: g.exp(self.secretKeys[l]) for l in range(1, k + 1)}
        self.fullPublicKey = getCiphPGroup(pGroup, 1).product(g, self.polynomialInExponent[0])
        self.nizkp = ProofDirectory()
        self.nizkp["version"] = params.version.encode()
        self.nizkp["type"] = b"mixing"
        self.nizkp["auxsid"] = params.auxsid.encode()
        self.nizkp["width"] = str(width).encode()
        self.nizkp["FullPublicKey.bt"] = self.fullPublicKey.toByteTree().to_bytes()
        self.nizkp["proofs/PolynomialInExponent.bt"] = \
            ByteTreeContainer(*[c.toByteTree() for c in self.polynomialInExponent]).to_bytes()

    def _party_source(self, l: int, what: str):
        """Every party draws from its own stream (seeded from the dealer's source, in party order)."""
        prg = PRGHeuristic()
        prg.setSeed(self.randomSource.getBytes(prg.minNoSeedBytes()))
        return prg

    # -- mixnet/MixNetElGamalSession: the first `threshold` parties shuffle in turn
    def shuffle(self, ciphertexts):
        self.nizkp["Ciphertexts.bt"] = ciphertexts.toByteTree().to_buffer()
        active = self.threshold
        self.nizkp["proofs/activethreshold"] = str(active).encode()
        inp, owned = ciphertexts, False
        for l in range(1, active + 1):
            session = ShufflerSession(self.pGroup, self.fullPublicKey, self.params, self._party_source(l, "shuffle"))
            proof, out = session.shuffle(self.width, inp, keep_output=True)
            # the last shuffler's output is the output of the mixing phase (MixNetElGamalSession.LSfile)
            self.nizkp["ShuffledCiphertexts.bt" if l == active else ProofDirectory.Lfile(l)] = proof.output
            self.nizkp[ProofDirectory.PCfile(l)] = proof.permutationCommitment
            self.nizkp[ProofDirectory.PoSCfile(l)] = proof.commitment
            self.nizkp[ProofDirectory.PoSRfile(l)] = proof.reply
            if owned:
                inp.free()
            inp, owned = out, True
        return inp

    # -- elgamal/DistrElGamalSession.java:361-545, every party's part
    def decrypt(self, ciphertexts):
        k, t, p = self.k, self.threshold, self.params
        g = self.pGroup.getg()
        u = ciphertexts.project(0)
        f = {l: eg.decryptionFactors(u, self.secretKeys[l], k) for l in range(1, k + 1)}
        for l in range(1, k + 1):
            self.nizkp[ProofDirectory.DFfile(l)] = f[l].toByteTree().to_buffer()
        correct = [False] + [True] * k
        combined = eg.combineDecryptionFactors(f, correct, k, t)
        challenger = ShufflerSession(self.pGroup, self.fullPublicKey, p, None).challenger
        seedData = _decryption_seed_data(g, ciphertexts, self.polynomialInExponent, f, k)
        prgSeed = challenger.challenge(seedData, 8 * PRGHeuristic().minNoSeedBytes(), p.rbitlen)
        parties = {}
        for l in range(1, k + 1):
            E = eg.DistrElGamalSessionBasic(l, k, t, p.ebitlenro, p.rbitlen, PRGHeuristic())
            E.setInstance(g, u, self.publicKeys, f, self.secretKeys[l], self.polynomialInExponent[0], combined)
            E.setBatchVector(prgSeed)
            E.batchInput()
            self.nizkp[ProofDirectory.DFCfile(l)] = E.commit(self._party_source(l, "decrypt")).to_bytes()
            parties[l] = E
        # every party reads the others' commitments; the challenge binds all of them
        E1 = parties[1]
        for l in range(2, k + 1):
            E1.setCommitment(l, self.nizkp[ProofDirectory.DFCfile(l)])
        challengeData = ByteTreeContainer(ByteTreeLeaf(prgSeed), E1.getCommitment())
        v = _to_positive(challenger.challenge(challengeData, p.vbitlenro, p.rbitlen))
        for l in range(1, k + 1):
            self.nizkp[ProofDirectory.DFRfile(l)] = parties[l].reply(v).to_bytes()
            parties[l].free()
        self.nizkp["proofs/CorrectIndices.bt"] = booleanArrayToByteTree(correct).to_bytes()
        plaintexts = ciphertexts.project(1).mul(combined)
        combined.free()
        for l in f:
            f[l].free()
        self.nizkp["Plaintexts.bt"] = plaintexts.toByteTree().to_buffer()
        return plaintexts

    def run(self, ciphertexts):
        shuffled = self.shuffle(ciphertexts)
        plaintexts = self.decrypt(shuffled)
        shuffled.free()
        return plaintexts


def _decryption_seed_data(g, ciphertexts, polynomialInExponent, f: Dict[int, object], k: int):
    """elgamal/DistrElGamalSession.java:433-462 = MixNetElGamalVerifyFiatShamirSession.java:1586-1600."""
    btIn = ByteTreeContainer(g.toByteTree(), ciphertexts.toByteTree())
    pkBT = ByteTreeContainer(*[c.toByteTree() for c in polynomialInExponent])
    dfBT = ByteTreeContainer(*[f[l].toByteTree() for l in range(1, k + 1)])
    return ByteTreeContainer(btIn, ByteTreeContainer(pkBT, dfBT))


# ---------------------------------------------------------------- vmnv
class MixNetElGamalVerifyFiatShamirSession:
    """mixnet/MixNetElGamalVerifyFiatShamirSession.java: verification of a proof of type "mixing" without
    pre-computation (verify:1318-1668).  `verify` returns a report; conditions under which the reference stops
    with an error raise VerificationError."""

    def __init__(self, pGroup, params: SessionParams, k: int, threshold: int, expectedAuxsid: Optional[str] = None,
                 expectedWidth: Optional[int] = None):
        """`expectedAuxsid`, `expectedWidth`: the `-auxsid` / `-width` options of vmnv; None accepts whatever the
        proof directory names."""
        self.pGroup, self.params, self.k, self.threshold = pGroup, params, k, threshold
        self.expectedAuxsid, self.expectedWidth = expectedAuxsid, expectedWidth
        self.report: Dict[str, object] = {}

    def _file(self, nizkp: ProofDirectory, name: str) -> bytes:
        if name not in nizkp:
            raise VerificationError("Can not find %s in proof directory!" % name)
        return nizkp[name]

    def _readArray(self, size: int, pGroup, data: bytes, name: str):
        try:
            return pGroup.toElementArray(size, ByteTreeReader(data))
        except (ArithmFormatException, EIOException) as e:
            raise VerificationError("Unable to read array %s! (%s)" % (name, e))

    def verify(self, nizkp: ProofDirectory) -> Dict[str, object]:
        self._spec = None
        try:
            return self._verify(nizkp)
        except (EIOException, ArithmFormatException, ValueError, UnicodeDecodeError) as e:
            # a malformed file outside the places where the reference substitutes trivial values is fail-stop
            # (mixnet/MixNetElGamalVerifyFiatShamirSession.java: `failStop`), never a stray parser exception
            raise VerificationError("Malformed proof directory: %s" % e)
        finally:
            if self._spec is not None:   # a fail-stop condition was met while the speculative hash was running
                self._spec.abandon()
                self._spec = None

    def _verify(self, nizkp: ProofDirectory) -> Dict[str, object]:
        p, k, threshold, G = self.params, self.k, self.threshold, self.pGroup
        rep = self.report = {"shuffles": {}, "decryption": None}
        if self._file(nizkp, "version").decode() != p.version:
            raise VerificationError("Mismatching versions!")
        if self._file(nizkp, "type").decode() != "mixing":
            raise VerificationError("Unsupported proof type")
        # determineAuxsid :369-395: the identifier is read from the proof, validated, compared with the expected
        # one if there is one, and enters the global prefix of every random-oracle call (setGlobalPrefix :158-189)
        auxsid = self._file(nizkp, "auxsid").decode("ascii", errors="replace")
        if not validateSid(auxsid):
            raise VerificationError("Can not read auxsid from file!")
        if self.expectedAuxsid is not None and auxsid != self.expectedAuxsid:
            raise VerificationError("The given auxiliary session identifier does not match the one in the proof!")
        p = dataclasses.replace(p, auxsid=auxsid)
        # determineWidth (:404-440): the number of ciphertexts shuffled in parallel; the keys of the directory are
        # the basic ones and are widened where they are used (elgamal/ProtocolElGamal.java:769-800)
        try:
            width = _parse_int(self._file(nizkp, "width"))
        except ValueError:
            raise VerificationError("Can not parse width given in file!")
        if width < 1 or (self.expectedWidth is not None and width != self.expectedWidth):
            raise VerificationError("Mismatching or invalid width!")
        ciphPGroup = getCiphPGroup(G, width)
        # readFullPKey :195-226
        try:
            fullPKey = getCiphPGroup(G, 1).toElement(ByteTreeReader(self._file(nizkp, "FullPublicKey.bt")))
        except (ArithmFormatException, EIOException):
            raise VerificationError("Could not read full El Gamal public key from file!")
        if not fullPKey.project(0).equals(G.getg()):
            raise VerificationError("Basic public key is not the standard generator!")
        # readMixServerPKeys :228-266
        try:
            btr = ByteTreeReader(self._file(nizkp, "proofs/PolynomialInExponent.bt"))
            if btr.isLeaf() or btr.getRemaining() != threshold:
                raise EIOException("degree")
            coeffs = [G.toElement(btr.getNextChild()) for _ in range(threshold)]
        except (ArithmFormatException, EIOException):
            raise VerificationError("Unable to read polynomial in exponent from file!")
        pkeys = {l: _evaluate_in_exponent(coeffs, l) for l in range(1, k + 1)}
        if not fullPKey.project(1).equals(coeffs[0]):
            raise VerificationError("Mismatching public keys!")
        session = ShufflerSession(G, fullPKey, p, None)
        challenger = session.challenger
        try:
            active = _parse_int(self._file(nizkp, "proofs/activethreshold"))
        except ValueError:
            raise VerificationError("Can not parse active threshold given in file!")
        if active > k or active < threshold:
            raise VerificationError("Active threshold out of range!")
        # readCiphertexts
        raw = self._file(nizkp, "Ciphertexts.bt")
        try:
            # Ciphertexts.bt = node(u, v); for width > 1 each of them is a node of `width` arrays; an array over a
            # curve group is node(x leaves, y leaves): descend to the first array of leaves
            r = ByteTreeReader(raw).getNextChild()
            for _ in range((1 if width > 1 else 0) + (1 if G.is_curve else 0)):
                r = r.getNextChild()
            size = r.getRemaining()
        except EIOException:
            raise VerificationError("Unable to read ciphertexts!")
        ciphertexts = self._readArray(size, ciphPGroup, raw, "Ciphertexts.bt")
        # ---- shuffles :1403-1520
        generators = session.deriveGenerators(size)
        inp, valid = ciphertexts, 0
        # The seed of the decryption proof is RO(node(node(g, L_active), node(node(pk), node(f_1..f_k)))) (:1586-1600):
        # 190 MB of SHA-256 at N = 10^5 that depend on files only.  It is hashed on a worker thread WHILE the
        # shuffles are verified, speculating that the last shuffle is valid and that the files are canonical;
        # the speculation is checked below and the hash redone in place if it does not hold.
        lastName = ProofDirectory.Lfile(active)
        if lastName not in nizkp:
            lastName = "ShuffledCiphertexts.bt"
        dfNames = [ProofDirectory.DFfile(l) for l in range(1, k + 1)]
        spec = None
        if lastName in nizkp and all(nm in nizkp for nm in dfNames):
            spec = challenger.begin(8 * PRGHeuristic().minNoSeedBytes())
            spec.update(node_header(2))
            spec.update(node_header(2))
            G.getg().toByteTree().update(spec)
            spec.update(nizkp[lastName])
            spec.update(node_header(2))
            ByteTreeContainer(*[c.toByteTree() for c in coeffs]).update(spec)
            spec.update(node_header(k))
            for nm in dfNames:
                spec.update(nizkp[nm])
            self._spec = spec
        for l in range(1, active + 1):
            name = ProofDirectory.Lfile(l)
            if l == active and name not in nizkp:
                name = "ShuffledCiphertexts.bt"
            for need in (ProofDirectory.PCfile(l), ProofDirectory.PoSCfile(l), ProofDirectory.PoSRfile(l), name):
                self._file(nizkp, need)
            proof = ShuffleProof(nizkp[name], nizkp[ProofDirectory.PCfile(l)], nizkp[ProofDirectory.PoSCfile(l)],
                                 nizkp[ProofDirectory.PoSRfile(l)])
            # readArray(output) is fail-stop in the reference; an invalid PROOF replaces the output by the input
            parsed = self._readArray(size, ciphPGroup, proof.output, name)
            verdict, out = session.verify(width, inp, proof, generators=generators, output=parsed)
            rep["shuffles"][l] = verdict
            valid += 1 if verdict else 0
            if inp is not ciphertexts:
                inp.free()
            inp = out
        generators.free()
        rep["validProofs"] = valid
        rep["enoughValidProofs"] = valid >= threshold
        mixed = inp
        # ---- decryption :1535-1665
        try:
            flags = ByteTreeReader(self._file(nizkp, "proofs/CorrectIndices.bt")).readBooleans(k + 1)
        except EIOException:
            raise VerificationError("Failed to read indices of correct decryption factors!")
        correct = [bool(x) for x in flags]
        if sum(correct[1:]) < threshold:
            raise VerificationError("Too few correct decryption factors!")
        u = mixed.project(0)
        f = {l: self._readArray(size, ciphPGroup.project(0), self._file(nizkp, ProofDirectory.DFfile(l)),
                                ProofDirectory.DFfile(l)) for l in range(1, k + 1)}
        combined = eg.combineDecryptionFactors(f, correct, k, threshold)
        basic = eg.DistrElGamalSessionBasic(0, k, threshold, p.ebitlenro, p.rbitlen, PRGHeuristic())
        basic.setInstance(G.getg(), u, pkeys, f, None, fullPKey.project(1), combined)
        holds = spec is not None and rep["shuffles"].get(active) is True and \
            len(nizkp[lastName]) == mixed.toByteTree().total_bytes() and \
            all(len(nizkp[nm]) == f[l].toByteTree().total_bytes() for l, nm in zip(range(1, k + 1), dfNames))
        self._spec = None
        if holds:
            prgSeed = challenger.finish(spec)
        else:
            if spec is not None:
                spec.abandon()
            seedData = _decryption_seed_data(G.getg(), mixed, coeffs, f, k)
            prgSeed = challenger.challenge(seedData, 8 * PRGHeuristic().minNoSeedBytes(), p.rbitlen)
        basic.setBatchVector(prgSeed)
        basic.batchInput()
        basic.batchCombined()
        for l in range(1, k + 1):
            basic.setCommitment(l, self._file(nizkp, ProofDirectory.DFCfile(l)))
        challengeData = ByteTreeContainer(ByteTreeLeaf(prgSeed), basic.getCommitment())
        v = _to_positive(challenger.challenge(challengeData, p.vbitlenro, p.rbitlen))
        for l in range(1, k + 1):
            basic.setReply(l, self._file(nizkp, ProofDirectory.DFRfile(l)))
        basic.combine(correct)
        ok = basic.verifyCombined(v)
        basic.free()
        for l in f:
            f[l].free()
        rep["decryption"] = ok
        if not ok:
            combined.free()
            if mixed is not ciphertexts:
                mixed.free()
            ciphertexts.free()
            raise VerificationError("Verify combined proof of decryption... failed!")
        computed = mixed.project(1).mul(combined)
        combined.free()
        plain = self._readArray(size, computed.getPGroup(), self._file(nizkp, "Plaintexts.bt"), "Plaintexts.bt")
        match = plain.equals(computed)
        plain.free()
        computed.free()
        if mixed is not ciphertexts:
            mixed.free()
        ciphertexts.free()
        rep["plaintexts"] = match
        if not match:
            raise VerificationError("Plaintexts are incorrect!")
        rep["accepted"] = bool(rep["enoughValidProofs"])
        return rep
