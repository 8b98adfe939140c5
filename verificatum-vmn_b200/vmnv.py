"""A whole mix on the engine, and its universal verification: the array work of

    mixnet/MixNetElGamalSession.java:161-358 (pre-computation, shuffling by the first `activeThreshold` parties --
    plain or commitment-consistent --, decryption; sessions of type "mixing", "shuffling", "decryption"),
    elgamal/DistrElGamalSession.java:361-545 (decryption factors, their exchange, the batched proof),
    mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668 (what `vmnv` does with a proof directory of any of
    those types: verifyPoS / verifyPoSC + shrinkPermComm + verifyCCPoS per party, the verification of the decryption;
    2 x verifyPoS + decryption is BASELINE.json config 3)

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
from .hvzk import CCPoSW, PoSCTW
from .mixnet import (CommittedShuffler, SessionParams, ShuffleProof, ShufflerSession, getCiphPGroup,
                     getWidePublicKey, validateSid)


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
    def PoSCCfile(l: int) -> str:   # hvzk/PoSCTW.java:221-235
        return "proofs/PoSCCommitment%02d.bt" % l

    @staticmethod
    def PoSCRfile(l: int) -> str:
        return "proofs/PoSCReply%02d.bt" % l

    @staticmethod
    def CCPoSCfile(l: int) -> str:  # hvzk/CCPoSW.java:274-288
        return "proofs/CCPoSCommitment%02d.bt" % l

    @staticmethod
    def CCPoSRfile(l: int) -> str:
        return "proofs/CCPoSReply%02d.bt" % l

    @staticmethod
    def KLfile(l: int) -> str:      # mixnet/PermutationCommitment.java:240-242
        return "proofs/KeepList%02d.bt" % l

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
    """Keys + pre-computation / shuffling / decryption of one list of ciphertexts (mixnet/MixNetElGamalSession.java:
    precomp :161-187, shuffle :208-245, decrypt :268-325, mix :345-358); everything published goes to `self.nizkp`."""

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
        self.committed = None     # party -> CommittedShuffler after precomp()
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

    # -- mixnet/MixNetElGamalSession.java:161-187 over ShufflerElGamalSession.precomp (:534-672): every active party
    # commits to a permutation of `maxciph` generators, proves it (PoSC) and pre-computes its re-encryption factors
    def precomp(self, maxciph: int) -> None:
        active = self.threshold
        self.nizkp["proofs/maxciph"] = str(maxciph).encode()
        self.nizkp["proofs/activethreshold"] = str(active).encode()
        self.nizkp["type"] = b"shuffling"
        self.committed = {}
        for l in range(1, active + 1):
            session = ShufflerSession(self.pGroup, self.fullPublicKey, self.params, self._party_source(l, "precomp"))
            cs = CommittedShuffler(session, self.width, maxciph)
            pc, c, r = cs.precomp()
            self.nizkp[ProofDirectory.PCfile(l)] = pc
            self.nizkp[ProofDirectory.PoSCCfile(l)] = c
            self.nizkp[ProofDirectory.PoSCRfile(l)] = r
            self.committed[l] = cs

    # -- :208-245: the first `threshold` parties shuffle in turn; after a pre-computation the shuffles are the
    # commitment-consistent ones over commitments shrunk to the actual number of ciphertexts (:972-1033)
    def shuffle(self, ciphertexts):
        self.nizkp["Ciphertexts.bt"] = ciphertexts.toByteTree().to_buffer()
        active = self.threshold
        self.nizkp["proofs/activethreshold"] = str(active).encode()
        self.nizkp["type"] = b"shuffling"
        inp, owned = ciphertexts, False
        if self.committed is not None:
            for l in range(1, active + 1):
                self.nizkp[ProofDirectory.KLfile(l)] = self.committed[l].shrink(ciphertexts.size())
        for l in range(1, active + 1):
            if self.committed is not None:
                cs = self.committed[l]
                cs.session.randomSource = self._party_source(l, "shuffle")
                proof, out = cs.shuffle(inp, keep_output=True)
                self.nizkp[ProofDirectory.CCPoSCfile(l)] = proof.commitment
                self.nizkp[ProofDirectory.CCPoSRfile(l)] = proof.reply
                cs.permutationCommitment.free()
                for a in (cs.generators, cs.reencExponents, cs.reencFactors):
                    a.free()
            else:
                session = ShufflerSession(self.pGroup, self.fullPublicKey, self.params, self._party_source(l, "shuffle"))
                proof, out = session.shuffle(self.width, inp, keep_output=True)
                self.nizkp[ProofDirectory.PCfile(l)] = proof.permutationCommitment
                self.nizkp[ProofDirectory.PoSCfile(l)] = proof.commitment
                self.nizkp[ProofDirectory.PoSRfile(l)] = proof.reply
            # the last shuffler's output is the output of the session (MixNetElGamalSession.LSfile)
            self.nizkp["ShuffledCiphertexts.bt" if l == active else ProofDirectory.Lfile(l)] = proof.output
            if owned:
                inp.free()
            inp, owned = out, True
        self.committed = None
        return inp

    # -- elgamal/DistrElGamalSession.java:361-545, every party's part
    def decrypt(self, ciphertexts):
        k, t, p = self.k, self.threshold, self.params
        g = self.pGroup.getg()
        self.nizkp["proofs/activethreshold"] = str(t).encode()                      # DistrElGamalSession.java:347-350
        if "ShuffledCiphertexts.bt" in self.nizkp:   # after a shuffle: its output moves into the proofs (:294-303)
            self.nizkp["type"] = b"mixing"
            self.nizkp[ProofDirectory.Lfile(t)] = self.nizkp.pop("ShuffledCiphertexts.bt")
        else:
            self.nizkp["type"] = b"decryption"
            self.nizkp["Ciphertexts.bt"] = ciphertexts.toByteTree().to_buffer()
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

    def run(self, ciphertexts, mode: str = "mixing", maxciph: Optional[int] = None):
        """mode "mixing" (:345-358): shuffle, then decrypt; "shuffling"; "decryption".  `maxciph`: pre-compute for
        that many ciphertexts first.  Returns the plaintexts (the shuffled list for "shuffling")."""
        if mode not in ("mixing", "shuffling", "decryption"):
            raise ValueError(mode)
        if mode == "decryption":
            return self.decrypt(ciphertexts)
        if maxciph is not None:
            if maxciph < ciphertexts.size():
                raise ValueError("more ciphertexts than pre-computed for")
            self.precomp(maxciph)
        shuffled = self.shuffle(ciphertexts)
        if mode == "shuffling":
            return shuffled
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
    """mixnet/MixNetElGamalVerifyFiatShamirSession.java: verification of a proof of type "mixing", "shuffling" or
    "decryption", with or without pre-computation (verify:1318-1668).  `verify` returns a report; conditions under
    which the reference stops with an error raise VerificationError."""

    def __init__(self, pGroup, params: SessionParams, k: int, threshold: int, expectedAuxsid: Optional[str] = None,
                 expectedWidth: Optional[int] = None, expectedType: Optional[str] = None, dec: bool = True,
                 posc: bool = True, ccpos: bool = True):
        """`expectedAuxsid`, `expectedWidth`, `expectedType`: the `-auxsid` / `-width` / `-mix|-shuffle|-decrypt`
        options of vmnv; None accepts whatever the proof directory names (vmnv without `-width` compares with the
        width of the protocol info file, determineWidth :404-440: pass that width to get its default behaviour).  `dec`, `posc`, `ccpos`: what is verified
        (`-nodec`, `-noposc`, `-noccpos`; mixnet/SessionParams.java)."""
        self.pGroup, self.params, self.k, self.threshold = pGroup, params, k, threshold
        self.expectedAuxsid, self.expectedWidth, self.expectedType = expectedAuxsid, expectedWidth, expectedType
        self.dec, self.posc, self.ccpos = dec, posc, ccpos
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

    def _first_array_size(self, raw: bytes, width: int) -> int:
        """readArray(0, ...): the size is that of the first array of leaves.  A ciphertext array is node(u, v); for
        width > 1 each of them is a node of `width` arrays; an array over a curve group is node(x leaves, y leaves)."""
        try:
            r = ByteTreeReader(raw).getNextChild()
            for _ in range((1 if width > 1 else 0) + (1 if self.pGroup.is_curve else 0)):
                r = r.getNextChild()
            if r.isLeaf():
                raise EIOException("array expected")
            return r.getRemaining()
        except EIOException:
            raise VerificationError("Unable to read ciphertexts!")

    def _verify(self, nizkp: ProofDirectory) -> Dict[str, object]:
        p, k, threshold, G = self.params, self.k, self.threshold, self.pGroup
        # the scalar test vectors of `vmnv -t` (mixnet/MixNetElGamalVerifyFiatShamirTool.java:82-224), in the order the
        # reference prints them: (name, party or None, value)
        vectors = []
        rep = self.report = {"shuffles": {}, "poscs": {}, "decryption": None, "vectors": vectors}

        def record(name, value, party=None):
            vectors.append((name, party, bytes(value).hex() if isinstance(value, (bytes, bytearray, memoryview))
                            else str(value)))
        if self._file(nizkp, "version").decode() != p.version:
            raise VerificationError("Mismatching versions!")
        # determineType :329-358, determineSessionParams :984-1005
        typ = self._file(nizkp, "type").decode("ascii", errors="replace")
        if typ not in ("mixing", "shuffling", "decryption"):
            raise VerificationError("Unknown type of proof!")
        if self.expectedType is not None and typ != self.expectedType:
            raise VerificationError("Attempting to verify proof of %s, but proof is a proof of %s!" %
                                    (self.expectedType, typ))
        rep["type"] = typ
        # determineAuxsid :369-395: the identifier is read from the proof, validated, compared with the expected
        # one if there is one, and enters the global prefix of every random-oracle call (setGlobalPrefix :158-189)
        auxsid = self._file(nizkp, "auxsid").decode("ascii", errors="replace")
        if not validateSid(auxsid):
            raise VerificationError("Can not read auxsid from file!")
        if self.expectedAuxsid is not None and auxsid != self.expectedAuxsid:
            raise VerificationError("The given auxiliary session identifier does not match the one in the proof!")
        p = dataclasses.replace(p, auxsid=auxsid)
        for name, value in (("par.k", k), ("par.lambda", threshold), ("par.n_e", p.ebitlenro), ("par.n_r", p.rbitlen),
                            ("par.n_v", p.vbitlenro), ("par.s_Gq", p.pGroupString), ("par.version", p.version)):
            record(name, value)
        dec, posc, ccpos = self.dec, self.posc, self.ccpos
        if typ == "shuffling":
            dec = False
        elif typ == "decryption":
            posc = ccpos = False
        # determineWidth (:404-440): the number of ciphertexts shuffled in parallel; the keys of the directory are
        # the basic ones and are widened where they are used (elgamal/ProtocolElGamal.java:769-800)
        width = 1
        if ccpos or dec:
            try:
                width = _parse_int(self._file(nizkp, "width"))
            except ValueError:
                raise VerificationError("Can not parse width given in file!")
            if width < 1 or (self.expectedWidth is not None and width != self.expectedWidth):
                raise VerificationError("Mismatching or invalid width!")
            record("par.omega", width)
        ciphPGroup = getCiphPGroup(G, width)
        # readFullPKey :195-226
        try:
            fullPKey = getCiphPGroup(G, 1).toElement(ByteTreeReader(self._file(nizkp, "FullPublicKey.bt")))
        except (ArithmFormatException, EIOException):
            raise VerificationError("Could not read full El Gamal public key from file!")
        if not fullPKey.project(0).equals(G.getg()):
            raise VerificationError("Basic public key is not the standard generator!")
        coeffs = pkeys = None
        if dec:   # readMixServerPKeys :228-266
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
        record("par.sid", p.sid)
        record("der.rho", challenger.globalPrefix)
        precomp = "proofs/maxciph" in nizkp                                                           # :946-948
        try:
            active = _parse_int(self._file(nizkp, "proofs/activethreshold"))
        except ValueError:
            raise VerificationError("Can not parse active threshold given in file!")
        if active > k or active < threshold:
            raise VerificationError("Active threshold out of range!")
        record("par.lambda", active)
        # readCiphertexts :1017-1046
        ciphertexts, ctName = None, None
        if ccpos or dec:
            if ccpos or typ == "decryption":
                ctName = "Ciphertexts.bt"
                self._file(nizkp, ctName)
            elif ProofDirectory.Lfile(active) in nizkp:
                ctName = ProofDirectory.Lfile(active)
            if ctName is not None:
                raw = nizkp[ctName]
                ciphertexts = self._readArray(self._first_array_size(raw, width), ciphPGroup, raw, ctName)
                if ciphertexts.size() == 0:
                    raise VerificationError("No ciphertexts!")
        # The seed of the decryption proof is RO(node(node(g, L), node(node(pk), node(f_1..f_k)))) (:1586-1600), L the
        # list that is decrypted: 190 MB of SHA-256 at N = 10^5 that depend on files only.  It is hashed on a worker
        # thread WHILE the shuffles are verified, speculating that L is the last output file (the last shuffle is
        # valid) and that the files are canonical; the speculation is checked below and the hash redone in place if
        # it does not hold.
        spec, lastName = None, None
        dfNames = [ProofDirectory.DFfile(l) for l in range(1, k + 1)]
        if dec:
            lastName = ctName
            if ccpos:
                lastName = ProofDirectory.Lfile(active)
                if lastName not in nizkp:
                    lastName = "ShuffledCiphertexts.bt"
            if lastName is not None and lastName in nizkp and all(nm in nizkp for nm in dfNames):
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
        mixed, mixedName = ciphertexts, ctName
        # ---- shuffles :1378-1530
        if posc or ccpos:
            if precomp:                                                                               # getMaxciph :541-548
                try:
                    maxciph = _parse_int(self._file(nizkp, "proofs/maxciph"))
                except ValueError:
                    raise VerificationError("Can not parse maxciph file!")
                record("par.N_0", maxciph)
                if maxciph < 1:
                    raise VerificationError("Invalid maxciph!")
                # every party that counts published a permutation commitment of maxciph elements (else
                # readPermutationCommitment :626-641 stops), maxciph leaves of at least a coordinate's length each:
                # a maxciph no file of the directory can answer to is refused before that many generators are derived
                if maxciph > max(len(v) for v in nizkp.values()) // (5 + (G.p.bit_length() + 7) // 8):
                    raise VerificationError("maxciph exceeds what the proof directory can hold!")
            else:
                if ciphertexts is None:
                    raise VerificationError("No ciphertexts!")
                maxciph = ciphertexts.size()
            generators = session.deriveGenerators(maxciph)                                            # :556-576
            shrunkGenerators = None
            if ccpos and precomp:                                                                     # :1059-1068
                if ciphertexts.size() > maxciph:
                    raise VerificationError("Too few generators have been derived!")
                shrunkGenerators = generators.copyOfRange(0, ciphertexts.size())
            widePublicKey = getWidePublicKey(fullPKey, width)
            inp, inpName, valid = ciphertexts, ctName, 0
            for l in range(1, active + 1):
                verdict = True
                pcName = ProofDirectory.PCfile(l)
                if not ((posc and precomp and not ccpos and pcName in nizkp)                          # getPoSCActive :958
                        or ProofDirectory.CCPoSCfile(l) in nizkp or ProofDirectory.PoSCfile(l) in nizkp):   # :972-976
                    continue
                self._file(nizkp, pcName)
                pc = None
                if precomp or not ccpos:   # readPermutationCommitment :626-641 (else PoSTW.verify parses it below)
                    pc = self._readArray(maxciph, G, nizkp[pcName], pcName)
                if posc and precomp:                                                                  # verifyPoSC :652-705
                    V = PoSCTW(p.vbitlenro, p.ebitlenro, p.rbitlen, session.prg, None, challenger)
                    ok = V.verify(G.getg(), generators, pc, self._file(nizkp, ProofDirectory.PoSCCfile(l)),
                                  self._file(nizkp, ProofDirectory.PoSCRfile(l)))
                    record("PoSC.s", V.testVector["s"], l)
                    record("PoSC.v", V.testVector["v"], l)
                    rep["poscs"][l] = ok
                    if not ok:
                        verdict = False
                        pc.free()
                        pc = generators.copyOfRange(0, maxciph)
                if ccpos:
                    size = inp.size()
                    name = ProofDirectory.Lfile(l)
                    if l == active and name not in nizkp:
                        name = "ShuffledCiphertexts.bt"
                    # readArray(output) is fail-stop in the reference; an invalid PROOF replaces the output by the input
                    output = self._readArray(size, ciphPGroup, self._file(nizkp, name), name)
                    if precomp:
                        # shrinkPermComm :714-745
                        try:
                            keep = ByteTreeReader(self._file(nizkp, ProofDirectory.KLfile(l))).readBooleans(maxciph)
                        except EIOException:
                            raise VerificationError("Unable to open keeplist of Party %d!" % l)
                        if int(keep.sum()) != size:
                            raise VerificationError("Wrong number of true elements in keep list of Party %d!" % l)
                        shrunk = pc.extract(keep)
                        pc.free()
                        pc = None
                        V = CCPoSW(p.vbitlenro, p.ebitlenro, p.rbitlen, session.prg, None, challenger)   # verifyCCPoS :757-841
                        ok = V.verify(G.getg(), shrunkGenerators, shrunk, widePublicKey, inp, output,
                                      self._file(nizkp, ProofDirectory.CCPoSCfile(l)),
                                      self._file(nizkp, ProofDirectory.CCPoSRfile(l)))
                        shrunk.free()
                        record("CCPoS.s", V.testVector["s"], l)
                        record("CCPoS.v", V.testVector["v"], l)
                        verdict = verdict and ok
                        if not verdict:
                            output.free()
                            output = inp.copyOfRange(0, size)
                    else:
                        for need in (ProofDirectory.PoSCfile(l), ProofDirectory.PoSRfile(l)):
                            self._file(nizkp, need)
                        proof = ShuffleProof(nizkp[name], nizkp[pcName], nizkp[ProofDirectory.PoSCfile(l)],
                                             nizkp[ProofDirectory.PoSRfile(l)])
                        verdict, output = session.verify(width, inp, proof, generators=generators, output=output)
                        if not session.last_u_parsed:   # readPermutationCommitment :626-641 is fail-stop
                            output.free()
                            raise VerificationError("Unable to read array %s!" % pcName)
                        record("PoS.s", session.last_test_vector["s"], l)
                        record("PoS.v", session.last_test_vector["v"], l)
                    if inp is not ciphertexts:
                        inp.free()
                    inp = output
                    if verdict:
                        inpName = name
                if pc is not None:
                    pc.free()
                rep["shuffles"][l] = verdict
                valid += 1 if verdict else 0
            generators.free()
            if shrunkGenerators is not None:
                shrunkGenerators.free()
            rep["validProofs"] = valid
            rep["enoughValidProofs"] = valid >= threshold
            if dec:
                mixed, mixedName = inp, inpName
            elif inp is not None and inp is not ciphertexts:
                inp.free()
        if not dec:
            if ciphertexts is not None:
                ciphertexts.free()
            rep["accepted"] = bool(rep.get("enoughValidProofs", True))
            return rep
        if mixed is None:
            raise VerificationError("No ciphertexts to decrypt!")
        size = mixed.size()
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
        holds = spec is not None and mixedName == lastName and \
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
        record("Dec.s", prgSeed)
        basic.setBatchVector(prgSeed)
        basic.batchInput()
        basic.batchCombined()
        for l in range(1, k + 1):
            basic.setCommitment(l, self._file(nizkp, ProofDirectory.DFCfile(l)))
        challengeData = ByteTreeContainer(ByteTreeLeaf(prgSeed), basic.getCommitment())
        v = _to_positive(challenger.challenge(challengeData, p.vbitlenro, p.rbitlen))
        record("Dec.v", v)
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
        rep["accepted"] = bool(rep.get("enoughValidProofs", True))
        return rep
