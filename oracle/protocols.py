"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement on Python integers of the protocol algebra of the hot path, following the
reference line by line:

  PoSBasicTW               hvzk/PoSBasicTW.java:407-482 (computeAF, precompute), :533-700 (commit),
                           :780-823 (setCommitment), :856-888 (reply), :1000-1066 (verify)
  PoSCBasicTW              hvzk/PoSCBasicTW.java:363-529, :607-636, :646-727
  CCPoSBasicW              hvzk/CCPoSBasicW.java:344-396, :462-485, :493-506, :519-584
  fiat_shamir_pos_*        hvzk/PoSTW.java:95-165 (prove), :177-260 (verify); hvzk/ChallengerRO.java:96-116
  shuffle                  mixnet/ShufflerElGamalSession.java:362-433, :250-300
  global_prefix            elgamal/ProtocolElGamal.java:659-683
  independent_generators   distr/IndependentGeneratorsRO.java:110-130
  demo_ciphertexts         elgamal/ProtocolElGamalInterfaceRaw.java:99-130
  decryption-factor proof  elgamal/DistrElGamalSessionBasic.java:513-540, :595-727

PARITY UNPINNED against a Java run (no JVM, no golden vectors in the reference).
Pinned by: tests/golden (PRG / RO / byte-tree KATs and the in-tree ModPGroup fixture), exact
integer arithmetic, and accept / reject behaviour (the only thing the reference's own
hvzk/TestPoSCBasicTW.java:147-163 asserts).
"""
from __future__ import annotations

from . import arithm as ar
from . import bytetree as bt
from .crypto import PRGHeuristic, RandomOracle, SeededRandomSource

import hashlib


# ---------------------------------------------------------------- hashing glue
def challenge(hashname: str, global_prefix: bytes, data: bt.ByteTree, out_bits: int) -> bytes:
    """hvzk/ChallengerRO.java:96-116."""
    d = RandomOracle(hashname, out_bits).get_digest()
    d.update(global_prefix)
    data.update(d)
    return d.digest()


def global_prefix(hashname, version, rosid, rbitlen, vbitlenro, ebitlenro, prg_string, pgroup_string, rohash_string):
    """elgamal/ProtocolElGamal.java:659-683."""
    t = bt.node(bt.string_leaf(version), bt.string_leaf(rosid), bt.int32_leaf(rbitlen), bt.int32_leaf(vbitlenro),
                bt.int32_leaf(ebitlenro), bt.string_leaf(prg_string), bt.string_leaf(pgroup_string),
                bt.string_leaf(rohash_string))
    return hashlib.new(hashname, t.to_bytes()).digest()


def independent_generators(G, hashname, prefix: bytes, sid: str, n: int, rbitlen: int):
    """distr/IndependentGeneratorsRO.java:110-130."""
    prg = PRGHeuristic(hashname)
    d = RandomOracle(hashname, 8 * prg.min_no_seed_bytes()).get_digest()
    d.update(prefix)
    d.update(bt.string_leaf(sid).to_bytes())
    prg.set_seed(d.digest())
    return ar.group_random_array(G, n, prg, rbitlen)


def demo_ciphertexts(G, pk, n: int, rs):
    """elgamal/ProtocolElGamalInterfaceRaw.java:99-130 (width 1)."""
    g, y = pk
    m = ar.group_random_array(G, n, rs, 10)
    r = ar.ring_random_array(G, n, rs, 20)
    u = ar.g_exp(G, g, r)
    t = ar.g_exp(G, y, r)
    return (u, ar.g_mul(G, t, m))


def batch_vector(hashname: str, seed: bytes, n: int, ebitlen: int):
    """setBatchVector (hvzk/PoSBasicTW.java:533-538)."""
    prg = PRGHeuristic(hashname)
    prg.set_seed(seed)
    return ar.lia_random(n, ebitlen, prg)


def _ring_shape(pk):
    """Shape of ciphPRing = pkey.project(0).getPGroup().getPRing() (None for Z_q)."""
    return tuple(None for _ in pk[0]) if isinstance(pk[0], tuple) else None


def _ring_random(G, shape, rs, rbitlen):
    if isinstance(shape, tuple):
        return tuple(_ring_random(G, s, rs, rbitlen) for s in shape)
    return ar.ring_random_element(G, rs, rbitlen)


def _rneg(G, x):
    return ar.gmap(lambda v: (-v) % G.q, x)


def _rmuladd(G, a, v, b):
    """a*v + b on ring elements or arrays."""
    def one(x, y):
        if isinstance(x, list):
            return [(s * v + t) % G.q for s, t in zip(x, y)]
        return (x * v + y) % G.q
    return ar.gmap(one, a, b)


# ---------------------------------------------------------------- PoSBasicTW
class PoSBasicTW:
    def __init__(self, G, vbitlen, ebitlen, rbitlen, prg_hash, rs):
        self.G, self.vbitlen, self.ebitlen, self.rbitlen, self.prg_hash, self.rs = G, vbitlen, ebitlen, rbitlen, prg_hash, rs

    # :436-482
    def precompute(self, g, h, pi=None):
        G = self.G
        self.g, self.h, self.size = g, h, len(h)
        if pi is None:
            return
        self.pi = pi
        self.r = ar.ring_random_array(G, self.size, self.rs, self.rbitlen)
        self.u = ar.permute(ar.g_mul(G, h, ar.g_exp(G, g, self.r)), pi)
        self.alpha = ar.ring_random_element(G, self.rs, self.rbitlen)
        # pField.toElementArray(epsilonIntegers) (:473): field elements, i.e. reduced mod q
        self.epsilon = [v % G.q for v in ar.lia_random(self.size, self.ebitlen + self.vbitlen + self.rbitlen, self.rs)]
        self.Ap = ar.g_mul(G, ar.g_exp(G, g, self.alpha), ar.g_exp_prod(G, h, self.epsilon))

    def set_instance(self, pkey, w, wp, s=None):
        self.pkey, self.w, self.wp, self.s = pkey, w, wp, s

    def set_batch_vector(self, seed: bytes):
        self.e = batch_vector(self.prg_hash, seed, self.size, self.ebitlen)

    # :407-410
    def compute_AF(self):
        self.A = ar.g_exp_prod(self.G, self.u, self.e)
        self.F = ar.g_exp_prod(self.G, self.w, self.e)

    # :546-700
    def commit(self, seed: bytes) -> bt.ByteTree:
        G, g, h = self.G, self.g, self.h
        self.set_batch_vector(seed)
        self.ipe = ar.permute(self.e, ar.perm_inv(self.pi))
        h0 = h[0]
        self.b = ar.ring_random_array(G, self.size, self.rs, self.rbitlen)
        x, self.d = ar.r_rec_lin(G, self.b, self.ipe)
        y = ar.r_prods(G, self.ipe)
        self.B = ar.g_mul(G, ar.g_exp(G, g, x), ar.g_exp(G, h0, y))
        self.beta = ar.ring_random_array(G, self.size, self.rs, self.rbitlen)
        xp = [0] + x[:-1]
        yp = [1] + y[:-1]
        e1 = [(bb + a * c) % G.q for bb, a, c in zip(self.beta, xp, self.epsilon)]
        e2 = [a * c % G.q for a, c in zip(yp, self.epsilon)]
        self.Bp = ar.g_mul(G, ar.g_exp(G, g, e1), ar.g_exp(G, h0, e2))
        self.gamma = ar.ring_random_element(G, self.rs, self.rbitlen)
        self.Cp = ar.g_exp(G, g, self.gamma)
        self.delta = ar.ring_random_element(G, self.rs, self.rbitlen)
        self.Dp = ar.g_exp(G, g, self.delta)
        self.phi = _ring_random(G, _ring_shape(self.pkey), self.rs, self.rbitlen)
        self.Fp = ar.g_mul(G, ar.g_exp(G, self.pkey, _rneg(G, self.phi)), ar.g_exp_prod(G, self.wp, self.epsilon))
        return self.commitment_tree()

    def commitment_tree(self) -> bt.ByteTree:
        G = self.G
        return bt.node(ar.array_tree(G, self.B), ar.elem_tree(G, self.Ap), ar.array_tree(G, self.Bp),
                       ar.elem_tree(G, self.Cp), ar.elem_tree(G, self.Dp), ar.elem_tree(G, self.Fp))

    # :780-823
    def set_commitment(self, t: bt.ByteTree) -> bt.ByteTree:
        G = self.G
        try:
            c = bt.first(t, 6)   # btr.getNextChild() six times (:787-805); a seventh child is never looked at
            self.B = ar.parse_array(G, c[0], self.size)
            self.Ap = ar.parse_elem(G, c[1])
            self.Bp = ar.parse_array(G, c[2], self.size)
            self.Cp = ar.parse_elem(G, c[3])
            self.Dp = ar.parse_elem(G, c[4])
            self.Fp = ar.parse_elem(G, c[5], self.pkey)
        except (ar.FormatError, bt.EIOError):
            self.B = [G.one] * self.size
            self.Bp = [G.one] * self.size
            self.Ap = self.Cp = self.Dp = G.one
            self.Fp = ar.gmap(lambda _: G.one, self.pkey)
        return self.commitment_tree()

    def set_challenge(self, v: int):
        assert 0 <= v and v.bit_length() <= self.vbitlen, "Malformed challenge!"
        self.v = v % self.G.q

    # :856-888
    def reply(self, v: int) -> bt.ByteTree:
        G = self.G
        self.set_challenge(v)
        v = self.v
        a = ar.r_inner(G, self.r, self.ipe)
        c = sum(self.r) % G.q
        f = ar.r_inner(G, self.s, self.e)
        self.k_A = (a * v + self.alpha) % G.q
        self.k_B = [(x * v + y) % G.q for x, y in zip(self.b, self.beta)]
        self.k_C = (c * v + self.gamma) % G.q
        self.k_D = (self.d * v + self.delta) % G.q
        self.k_E = [(x * v + y) % G.q for x, y in zip(self.ipe, self.epsilon)]
        self.k_F = _rmuladd(G, f, v, self.phi)
        return self.reply_tree()

    def reply_tree(self) -> bt.ByteTree:
        G = self.G
        return bt.node(ar.ring_tree(G, self.k_A), ar.ring_array_tree(G, self.k_B), ar.ring_tree(G, self.k_C),
                       ar.ring_tree(G, self.k_D), ar.ring_array_tree(G, self.k_E), ar.ring_tree(G, self.k_F))

    # :970-990, :1000-1066
    def verify(self, t: bt.ByteTree) -> bool:
        G, g, h, u = self.G, self.g, self.h, self.u
        try:
            c = bt.first(t, 6)   # (:975-986)
            self.k_A = ar.parse_ring(G, c[0])
            self.k_B = ar.parse_ring_array(G, c[1], self.size)
            self.k_C = ar.parse_ring(G, c[2])
            self.k_D = ar.parse_ring(G, c[3])
            self.k_E = ar.parse_ring_array(G, c[4], self.size)
            self.k_F = ar.parse_ring(G, c[5], _ring_shape(self.pkey))
        except (ar.FormatError, bt.EIOError):
            return False
        v = self.v
        h0 = h[0]
        C = ar.g_mul(G, ar.g_prod(G, u), ar.g_inv(G, ar.g_prod(G, h)))
        eprod = 1
        for x in self.e:
            eprod = eprod * x % G.q
        D = ar.g_mul(G, self.B[-1], ar.g_inv(G, ar.g_exp(G, h0, eprod)))
        vA = ar.g_mul(G, ar.g_exp(G, self.A, v), self.Ap) == ar.g_mul(G, ar.g_exp(G, g, self.k_A), ar.g_exp_prod(G, h, self.k_E))
        left = ar.g_mul(G, ar.g_exp(G, self.B, v), self.Bp)
        right = ar.g_mul(G, ar.g_exp(G, g, self.k_B), ar.g_exp(G, [h0] + self.B[:-1], self.k_E))
        vB = left == right
        vC = ar.g_mul(G, ar.g_exp(G, C, v), self.Cp) == ar.g_exp(G, g, self.k_C)
        vD = ar.g_mul(G, ar.g_exp(G, D, v), self.Dp) == ar.g_exp(G, g, self.k_D)
        lhsF = ar.g_mul(G, ar.g_exp(G, self.F, v), self.Fp)
        rhsF = ar.g_mul(G, ar.g_exp(G, self.pkey, _rneg(G, self.k_F)), ar.g_exp_prod(G, self.wp, self.k_E))
        vF = lhsF == rhsF
        self.verdicts = (vA, vB, vC, vD, vF)
        return all(self.verdicts)


# ---------------------------------------------------------------- Fiat-Shamir wrapper (hvzk/PoSTW.java)
class Params:
    def __init__(self, vbitlenro=256, ebitlenro=256, rbitlen=100, rohash="sha256", prghash="sha256",
                 version="3.1.0", sid="vmx", auxsid="default", pgroup_string=""):
        self.vbitlenro, self.ebitlenro, self.rbitlen = vbitlenro, ebitlenro, rbitlen
        self.rohash, self.prghash, self.version, self.pgroup_string = rohash, prghash, version, pgroup_string
        self.sid, self.auxsid = sid, auxsid

    @property
    def rosid(self) -> str:
        """mixnet/MixNetElGamalVerifyFiatShamirSession.java:160."""
        return self.sid + "." + self.auxsid

    def with_auxsid(self, auxsid: str) -> "Params":
        return Params(self.vbitlenro, self.ebitlenro, self.rbitlen, self.rohash, self.prghash, self.version, self.sid,
                      auxsid, self.pgroup_string)

    def prefix(self) -> bytes:
        names = {"sha256": "SHA-256", "sha384": "SHA-384", "sha512": "SHA-512"}
        return global_prefix(self.rohash, self.version, self.rosid, self.rbitlen, self.vbitlenro, self.ebitlenro,
                             "PRGHeuristic(%s)" % names[self.prghash], self.pgroup_string,
                             "HashfunctionHeuristic(%s)" % names[self.rohash])


def _seed_data(G, P, pkey, w, wp) -> bt.ByteTree:
    """hvzk/PoSTW.java:118-124."""
    return bt.node(ar.elem_tree(G, P.g), ar.array_tree(G, P.h), ar.array_tree(G, P.u), ar.elem_tree(G, pkey),
                   ar.array_tree(G, w), ar.array_tree(G, wp))


def shuffle_and_prove(G, params: Params, pkey, w, h, rs):
    """mixnet/ShufflerElGamalSession.java:400-414 + :273-289 + hvzk/PoSTW.java:95-165.
    Returns (output array, dict of the four published byte strings)."""
    n = ar.size_of(w)
    prefix = params.prefix()
    shape = _ring_shape(pkey)
    if isinstance(shape, tuple):
        s = tuple(ar.ring_random_array(G, n, rs, params.rbitlen) for _ in shape)
    else:
        s = ar.ring_random_array(G, n, rs, params.rbitlen)
    factors = ar.g_exp(G, pkey, s)
    pi = ar.permutation_random(n, rs, params.rbitlen)
    P = PoSBasicTW(G, params.vbitlenro, params.ebitlenro, params.rbitlen, params.prghash, rs)
    P.precompute(G.g, h, pi)
    wp = ar.permute(ar.g_mul(G, w, factors), ar.perm_inv(pi))
    P.set_instance(pkey, w, wp, s)
    seed = challenge(params.rohash, prefix, _seed_data(G, P, pkey, w, wp), 8 * PRGHeuristic(params.prghash).min_no_seed_bytes())
    commitment = P.commit(seed)
    cb = challenge(params.rohash, prefix, bt.node(bt.leaf(seed), commitment), params.vbitlenro)
    reply = P.reply(int.from_bytes(cb, "big"))
    return wp, {"output": ar.array_tree(G, wp).to_bytes(), "permutationCommitment": ar.array_tree(G, P.u).to_bytes(),
                "commitment": commitment.to_bytes(), "reply": reply.to_bytes()}


def verify_shuffle(G, params: Params, pkey, w, h, proof: dict, tv=None) -> bool:
    """mixnet/ShufflerElGamalSession.java:195-210,301-330 + hvzk/PoSTW.java:177-260.  `tv`: a dict that receives the
    seed ("s") and the challenge ("v"), the test vectors PoS.s / PoS.v of `vmnv -t`."""
    n = ar.size_of(w)
    prefix = params.prefix()
    try:
        wp = ar.parse_array(G, bt.read(proof["output"]), n, pkey)
    except (ar.FormatError, bt.EIOError):
        return False
    V = PoSBasicTW(G, params.vbitlenro, params.ebitlenro, params.rbitlen, params.prghash, None)
    V.precompute(G.g, h)
    V.set_instance(pkey, w, wp)
    try:
        V.u = ar.parse_array(G, bt.read(proof["permutationCommitment"]), n)
    except (ar.FormatError, bt.EIOError):
        V.u = list(h)
    seed = challenge(params.rohash, prefix, _seed_data(G, V, pkey, w, wp), 8 * PRGHeuristic(params.prghash).min_no_seed_bytes())
    V.set_batch_vector(seed)
    V.compute_AF()
    try:
        ctree = V.set_commitment(bt.read(proof["commitment"]))
    except bt.EIOError:
        ctree = V.set_commitment(bt.leaf(b""))
    cb = challenge(params.rohash, prefix, bt.node(bt.leaf(seed), ctree), params.vbitlenro)
    if tv is not None:
        tv["s"], tv["v"] = seed, int.from_bytes(cb, "big")
    V.set_challenge(int.from_bytes(cb, "big"))
    try:
        return V.verify(bt.read(proof["reply"]))
    except bt.EIOError:
        return False


# ---------------------------------------------------------------- PoSCBasicTW
class PoSCBasicTW:
    """hvzk/PoSCBasicTW.java:306-727: proof of a shuffle of commitments u = (h * g^r) permuted.
    Order in which the prover draws from its random source (:363-446): b, alpha, epsilon, beta,
    gamma, delta."""

    def __init__(self, G, vbitlen, ebitlen, rbitlen, prg_hash, rs):
        self.G, self.vbitlen, self.ebitlen, self.rbitlen, self.prg_hash, self.rs = G, vbitlen, ebitlen, rbitlen, prg_hash, rs

    def set_instance(self, g, h, u, r=None, pi=None):
        self.g, self.h, self.u, self.r, self.pi, self.size = g, h, u, r, pi, len(h)

    def set_batch_vector(self, seed: bytes):
        self.e = batch_vector(self.prg_hash, seed, self.size, self.ebitlen)

    # :363-446
    def commit(self, seed: bytes) -> bt.ByteTree:
        G, g, h = self.G, self.g, self.h
        self.set_batch_vector(seed)
        self.ipe = ar.permute(self.e, ar.perm_inv(self.pi))
        h0 = h[0]
        self.b = ar.ring_random_array(G, self.size, self.rs, self.rbitlen)
        x, self.d = ar.r_rec_lin(G, self.b, self.ipe)
        y = ar.r_prods(G, self.ipe)
        self.B = ar.g_mul(G, ar.g_exp(G, g, x), ar.g_exp(G, h0, y))
        self.alpha = ar.ring_random_element(G, self.rs, self.rbitlen)
        self.epsilon = [v % G.q for v in ar.lia_random(self.size, self.ebitlen + self.vbitlen + self.rbitlen, self.rs)]
        self.Ap = ar.g_mul(G, ar.g_exp(G, g, self.alpha), ar.g_exp_prod(G, h, self.epsilon))
        self.beta = ar.ring_random_array(G, self.size, self.rs, self.rbitlen)
        xp = [0] + x[:-1]
        yp = [1] + y[:-1]
        e1 = [(bb + a * c) % G.q for bb, a, c in zip(self.beta, xp, self.epsilon)]
        e2 = [a * c % G.q for a, c in zip(yp, self.epsilon)]
        self.Bp = ar.g_mul(G, ar.g_exp(G, g, e1), ar.g_exp(G, h0, e2))
        self.gamma = ar.ring_random_element(G, self.rs, self.rbitlen)
        self.Cp = ar.g_exp(G, g, self.gamma)
        self.delta = ar.ring_random_element(G, self.rs, self.rbitlen)
        self.Dp = ar.g_exp(G, g, self.delta)
        return self.commitment_tree()

    def commitment_tree(self) -> bt.ByteTree:
        G = self.G
        return bt.node(ar.array_tree(G, self.B), ar.elem_tree(G, self.Ap), ar.array_tree(G, self.Bp),
                       ar.elem_tree(G, self.Cp), ar.elem_tree(G, self.Dp))

    # :464-500
    def set_commitment(self, t: bt.ByteTree) -> bt.ByteTree:
        G = self.G
        try:
            if t.is_leaf() or len(t.children) < 5:
                raise ar.FormatError("commitment arity")
            c = t.children
            self.B = ar.parse_array(G, c[0], self.size)
            self.Ap = ar.parse_elem(G, c[1])
            self.Bp = ar.parse_array(G, c[2], self.size)
            self.Cp = ar.parse_elem(G, c[3])
            self.Dp = ar.parse_elem(G, c[4])
        except ar.FormatError:
            self.B = [G.one] * self.size
            self.Bp = [G.one] * self.size
            self.Ap = self.Cp = self.Dp = G.one
        return self.commitment_tree()

    def set_challenge(self, v: int):
        assert 0 <= v and v.bit_length() <= self.vbitlen, "Malformed challenge!"
        self.v = v % self.G.q

    # :607-636
    def reply(self, v: int) -> bt.ByteTree:
        G = self.G
        self.set_challenge(v)
        v = self.v
        a = ar.r_inner(G, self.r, self.ipe)
        c = sum(self.r) % G.q
        self.k_A = (a * v + self.alpha) % G.q
        self.k_B = [(x * v + y) % G.q for x, y in zip(self.b, self.beta)]
        self.k_C = (c * v + self.gamma) % G.q
        self.k_D = (self.d * v + self.delta) % G.q
        self.k_E = [(x * v + y) % G.q for x, y in zip(self.ipe, self.epsilon)]
        return bt.node(ar.ring_tree(G, self.k_A), ar.ring_array_tree(G, self.k_B), ar.ring_tree(G, self.k_C),
                       ar.ring_tree(G, self.k_D), ar.ring_array_tree(G, self.k_E))

    # :646-727  (checks are short-circuited in the reference; the verdict is their conjunction)
    def verify(self, t: bt.ByteTree) -> bool:
        G, g, h, u = self.G, self.g, self.h, self.u
        try:
            if t.is_leaf() or len(t.children) < 5:
                raise ar.FormatError("reply arity")
            c = t.children
            self.k_A = ar.parse_ring(G, c[0])
            self.k_B = ar.parse_ring_array(G, c[1], self.size)
            self.k_C = ar.parse_ring(G, c[2])
            self.k_D = ar.parse_ring(G, c[3])
            self.k_E = ar.parse_ring_array(G, c[4], self.size)
        except ar.FormatError:
            return False
        v, h0 = self.v, h[0]
        A = ar.g_exp_prod(G, u, self.e)
        C = ar.g_mul(G, ar.g_prod(G, u), ar.g_inv(G, ar.g_prod(G, h)))
        eprod = 1
        for x in self.e:
            eprod = eprod * x % G.q
        D = ar.g_mul(G, self.B[-1], ar.g_inv(G, ar.g_exp(G, h0, eprod)))
        vA = ar.g_mul(G, ar.g_exp(G, A, v), self.Ap) == ar.g_mul(G, ar.g_exp(G, g, self.k_A), ar.g_exp_prod(G, h, self.k_E))
        left = ar.g_mul(G, ar.g_exp(G, self.B, v), self.Bp)
        right = ar.g_mul(G, ar.g_exp(G, g, self.k_B), ar.g_exp(G, [h0] + self.B[:-1], self.k_E))
        vB = left == right
        vC = ar.g_mul(G, ar.g_exp(G, C, v), self.Cp) == ar.g_exp(G, g, self.k_C)
        vD = ar.g_mul(G, ar.g_exp(G, D, v), self.Dp) == ar.g_exp(G, g, self.k_D)
        self.verdicts = (vA, vB, vC, vD)
        return all(self.verdicts)


# ---------------------------------------------------------------- CCPoSBasicW
class CCPoSBasicW:
    """hvzk/CCPoSBasicW.java:268-584: commitment-consistent proof of a shuffle (the permutation is
    fixed by the commitment u; only multi-exponentiations)."""

    def __init__(self, G, vbitlen, ebitlen, rbitlen, prg_hash):
        self.G, self.vbitlen, self.ebitlen, self.rbitlen, self.prg_hash = G, vbitlen, ebitlen, rbitlen, prg_hash

    def set_instance(self, g, h, u, pkey, w, wp, r=None, pi=None, s=None):
        self.g, self.h, self.u, self.pkey, self.w, self.wp = g, h, u, pkey, w, wp
        self.r, self.pi, self.s, self.size = r, pi, s, len(h)

    def set_batch_vector(self, seed: bytes):
        self.e = batch_vector(self.prg_hash, seed, self.size, self.ebitlen)

    # :344-396
    def commit(self, seed: bytes, rs) -> bt.ByteTree:
        G = self.G
        self.set_batch_vector(seed)
        self.ipe = ar.permute(self.e, ar.perm_inv(self.pi))
        self.alpha = ar.ring_random_element(G, rs, self.rbitlen)
        self.epsilon = [v % G.q for v in ar.lia_random(self.size, self.ebitlen + self.vbitlen + self.rbitlen, rs)]
        self.Ap = ar.g_mul(G, ar.g_exp(G, self.g, self.alpha), ar.g_exp_prod(G, self.h, self.epsilon))
        self.beta = _ring_random(G, _ring_shape(self.pkey), rs, self.rbitlen)
        self.Bp = ar.g_mul(G, ar.g_exp(G, self.pkey, _rneg(G, self.beta)), ar.g_exp_prod(G, self.wp, self.epsilon))
        return self.commitment_tree()

    def commitment_tree(self) -> bt.ByteTree:
        return bt.node(ar.elem_tree(self.G, self.Ap), ar.elem_tree(self.G, self.Bp))

    # :407-430
    def set_commitment(self, t: bt.ByteTree) -> bt.ByteTree:
        G = self.G
        try:
            if t.is_leaf() or len(t.children) < 2:
                raise ar.FormatError("commitment arity")
            self.Ap = ar.parse_elem(G, t.children[0])
            self.Bp = ar.parse_elem(G, t.children[1], self.pkey)
        except ar.FormatError:
            self.Ap = G.one
            self.Bp = ar.gmap(lambda _: G.one, self.pkey)
        return self.commitment_tree()

    def set_challenge(self, v: int):
        assert 0 <= v and v.bit_length() <= self.vbitlen, "Malformed challenge!"
        self.v = v % self.G.q

    # :462-485
    def reply(self, v: int) -> bt.ByteTree:
        G = self.G
        self.set_challenge(v)
        v = self.v
        a = ar.r_inner(G, self.r, self.ipe)
        b = ar.r_inner(G, self.s, self.e)
        self.k_A = (a * v + self.alpha) % G.q
        self.k_B = _rmuladd(G, b, v, self.beta)
        self.k_E = [(x * v + y) % G.q for x, y in zip(self.ipe, self.epsilon)]
        return bt.node(ar.ring_tree(G, self.k_A), ar.ring_tree(G, self.k_B), ar.ring_array_tree(G, self.k_E))

    # :493-506 (raisedu == null)
    def compute_AB(self):
        self.A = ar.g_exp_prod(self.G, self.u, self.e)
        self.B = ar.g_exp_prod(self.G, self.w, self.e)

    # :519-584 (raisedExponent == null)
    def verify(self, t: bt.ByteTree) -> bool:
        G = self.G
        try:
            if t.is_leaf() or len(t.children) < 3:
                raise ar.FormatError("reply arity")
            self.k_A = ar.parse_ring(G, t.children[0])
            self.k_B = ar.parse_ring(G, t.children[1], _ring_shape(self.pkey))
            self.k_E = ar.parse_ring_array(G, t.children[2], self.size)
        except ar.FormatError:
            return False
        v = self.v
        vA = ar.g_mul(G, ar.g_exp(G, self.A, v), self.Ap) == ar.g_mul(G, ar.g_exp(G, self.g, self.k_A), ar.g_exp_prod(G, self.h, self.k_E))
        lhs = ar.g_mul(G, ar.g_exp(G, self.B, v), self.Bp)
        rhs = ar.g_mul(G, ar.g_exp(G, self.pkey, _rneg(G, self.k_B)), ar.g_exp_prod(G, self.wp, self.k_E))
        vB = lhs == rhs
        self.verdicts = (vA, vB)
        return vA and vB


# ---------------------------------------------------------------- decryption factors and their proof
# ODD_PRIME_TABLE (:199-215): the odd primes up to 1009
_ODD_PRIMES = [n for n in range(3, 1010, 2) if all(n % d for d in range(3, int(n ** 0.5) + 1, 2))]


def prime_log(number: int, prime: int) -> int:
    """elgamal/DistrElGamalSessionBasic.java:290-302: largest power of `prime` that is <= number."""
    a = b = 1
    while b <= number:
        a = b
        b *= prime
    return a


def prod_factor(q: int, k: int) -> int:
    """:304-325."""
    res, prime, i = 1, 2, 0
    while prime <= k:
        res *= prime_log(k, prime)
        prime = _ODD_PRIMES[i]
        i += 1
    return res * res % q


def modified_lagrange_coefficients(q: int, correct, k: int, threshold: int):
    """:327-452: small signed integers; correct[1..k]."""
    pf = prod_factor(q, k)
    out = []
    i = 1
    while len(out) < threshold and i <= k:
        if correct[i]:
            res, t, l = pf, 0, 1
            while t < threshold and l <= k:
                if correct[l]:
                    if l != i:
                        res = res * l % q
                        res = res * pow((l - i) % q, -1, q) % q
                    t += 1
                l += 1
            alt = res - q
            out.append(alt if abs(alt) < res else res)
        i += 1
    if len(out) < threshold:
        raise ValueError("Attempting to combine too few decryption factors!")
    return out


def decryption_factors(G, u, x: int, inverse_factor_scalar: int):
    """elgamal/DistrElGamalSession.java:377-385: f_i = u_i^{-x * c} for all first components."""
    ex = (-x * inverse_factor_scalar) % G.q
    return ar.g_exp(G, u, ex)


def combine_decryption_factors(G, factors: dict, correct, k: int, threshold: int):
    """:454-503: element-wise prod_j f_j^{lambda_j} with small signed integers."""
    ints = modified_lagrange_coefficients(G.q, correct, k, threshold)
    bases = [factors[i] for i in range(1, k + 1) if correct[i]][:threshold]
    out = None
    for b, lam in zip(bases, ints):
        t = ar.g_exp(G, b, lam % G.q)
        out = t if out is None else ar.g_mul(G, out, t)
    return out


class DistrElGamalSessionBasic:
    """elgamal/DistrElGamalSessionBasic.java:59: batched sigma proof that party l's decryption
    factors f_l = u^{-x_l c} are consistent with y_l = g^{x_l} (c = inverseFactor)."""

    def __init__(self, G, j, k, threshold, ebitlen, rbitlen, prg_hash, g, y: dict, u, x=None):
        self.G, self.j, self.k, self.threshold, self.ebitlen, self.rbitlen, self.prg_hash = G, j, k, threshold, ebitlen, rbitlen, prg_hash
        self.g, self.y, self.u, self.x = g, y, u, x
        self.inverse_factor = pow(prod_factor(G.q, k), -1, G.q)      # :243-249
        self.f, self.B, self.yp, self.Bp, self.k_x = {}, {}, {}, {}, {}
        self.verdicts = {l: True for l in range(1, k + 1)}

    def set_batch_vector(self, seed: bytes):                         # :513-518
        self.e = batch_vector(self.prg_hash, seed, ar.size_of(self.u), self.ebitlen)

    def batch_input(self):                                           # :524-526
        self.A = ar.g_exp_prod(self.G, self.u, self.e)

    def commit(self, rs) -> bt.ByteTree:                             # :534-540
        G = self.G
        self.r = ar.ring_random_element(G, rs, self.rbitlen)
        self.yp[self.j] = ar.g_exp(G, self.g, self.r)
        self.Bp[self.j] = ar.g_exp(G, self.A, self.r)
        return self.commitment_tree(self.j)

    def commitment_tree(self, l) -> bt.ByteTree:
        return bt.node(ar.elem_tree(self.G, self.yp[l]), ar.elem_tree(self.G, self.Bp[l]))

    def set_commitment(self, l, t: bt.ByteTree):                     # :549-565
        G = self.G
        try:
            if t.is_leaf() or len(t.children) < 2:
                raise ar.FormatError("commitment arity")
            self.yp[l] = ar.parse_elem(G, t.children[0])
            self.Bp[l] = ar.parse_elem(G, t.children[1], self.A)
        except ar.FormatError:
            self.verdicts[l] = False
            self.yp[l] = self.G.one
            self.Bp[l] = ar.gmap(lambda _: self.G.one, self.A)

    def reply(self, v: int) -> bt.ByteTree:                          # :595-598
        G = self.G
        self.k_x[self.j] = ((-self.x) * self.inverse_factor % G.q * (v % G.q) + self.r) % G.q
        return ar.ring_tree(G, self.k_x[self.j])

    def set_reply(self, l, t: bt.ByteTree):                          # :606-614
        try:
            self.k_x[l] = ar.parse_ring(self.G, t)
        except ar.FormatError:
            self.k_x[l] = 0
            self.verdicts[l] = False

    def batch(self, l):                                              # :707-709
        self.B[l] = ar.g_exp_prod(self.G, self.f[l], self.e)

    def verify(self, l, v: int) -> bool:                             # :718-727
        G = self.G
        if not self.verdicts[l]:
            return False
        pfev = v % G.q
        lhs1 = ar.g_mul(G, ar.g_exp(G, ar.g_inv(G, self.y[l]), self.inverse_factor * pfev % G.q), self.yp[l])
        ok1 = lhs1 == ar.g_exp(G, self.g, self.k_x[l])
        lhs2 = ar.g_mul(G, ar.g_exp(G, self.B[l], pfev), self.Bp[l])
        ok2 = lhs2 == ar.g_exp(G, self.A, self.k_x[l])
        return ok1 and ok2

    def combine(self, correct):                                      # :642-678
        G = self.G
        ints = modified_lagrange_coefficients(G.q, correct, self.k, self.threshold)
        exps = [i % G.q for i in ints]
        self.combinedyp = G.one
        self.combinedBp = ar.gmap(lambda _: G.one, self.A)
        self.combinedk_x = 0
        t, l = 0, 1
        while t < self.threshold and l <= self.k:
            if correct[l]:
                self.combinedyp = ar.g_mul(G, self.combinedyp, ar.g_exp(G, self.yp[l], exps[t]))
                self.combinedBp = ar.g_mul(G, self.combinedBp, ar.g_exp(G, self.Bp[l], exps[t]))
                self.combinedk_x = (self.combinedk_x + self.k_x[l] * exps[t]) % G.q
                t += 1
            l += 1

    def batch_combined(self, combinedf):                             # :683-685
        self.combinedB = ar.g_exp_prod(self.G, combinedf, self.e)

    def verify_combined(self, combinedy, v: int) -> bool:            # :693-700
        G = self.G
        pfev = v % G.q
        ok1 = ar.g_mul(G, ar.g_exp(G, ar.g_inv(G, combinedy), pfev), self.combinedyp) == ar.g_exp(G, self.g, self.combinedk_x)
        ok2 = ar.g_mul(G, ar.g_exp(G, self.combinedB, pfev), self.combinedBp) == ar.g_exp(G, self.A, self.combinedk_x)
        return ok1 and ok2


# ---------------------------------------------------------------- pre-computation and committed shuffle
def perm_shrink(table, size: int):
    """Permutation.shrink ([VCR-mem]; used at mixnet/PermutationCommitment.java:426): restriction to the
    first `size` inputs with images renumbered in increasing order."""
    img = list(table[:size])
    rank = {v: i for i, v in enumerate(sorted(img))}
    return [rank[v] for v in img]


def posc_prove(G, params: Params, g, h, u, r, pi, rs):
    """hvzk/PoSCTW.java:73-128."""
    prefix = params.prefix()
    P = PoSCBasicTW(G, params.vbitlenro, params.ebitlenro, params.rbitlen, params.prghash, rs)
    P.set_instance(g, h, u, r, pi)
    seed = challenge(params.rohash, prefix, bt.node(ar.elem_tree(G, g), ar.array_tree(G, h), ar.array_tree(G, u)),
                     8 * PRGHeuristic(params.prghash).min_no_seed_bytes())
    commitment = P.commit(seed)
    cb = challenge(params.rohash, prefix, bt.node(bt.leaf(seed), commitment), params.vbitlenro)
    reply = P.reply(int.from_bytes(cb, "big"))
    return commitment.to_bytes(), reply.to_bytes()


def posc_verify(G, params: Params, g, h, u, commitment: bytes, reply: bytes, tv=None) -> bool:
    """hvzk/PoSCTW.java:137-210; mixnet/MixNetElGamalVerifyFiatShamirSession.java:652-705."""
    prefix = params.prefix()
    V = PoSCBasicTW(G, params.vbitlenro, params.ebitlenro, params.rbitlen, params.prghash, None)
    V.set_instance(g, h, u)
    seed = challenge(params.rohash, prefix, bt.node(ar.elem_tree(G, g), ar.array_tree(G, h), ar.array_tree(G, u)),
                     8 * PRGHeuristic(params.prghash).min_no_seed_bytes())
    V.set_batch_vector(seed)
    try:
        ctree = V.set_commitment(bt.read(commitment))
    except bt.EIOError:
        ctree = V.set_commitment(bt.leaf(b""))
    cb = challenge(params.rohash, prefix, bt.node(bt.leaf(seed), ctree), params.vbitlenro)
    if tv is not None:
        tv["s"], tv["v"] = seed, int.from_bytes(cb, "big")
    V.set_challenge(int.from_bytes(cb, "big"))
    try:
        return V.verify(bt.read(reply))
    except bt.EIOError:
        return False


def _ccpos_seed(G, params, g, h, u, pkey, w, wp):
    return challenge(params.rohash, params.prefix(),
                     bt.node(ar.elem_tree(G, g), ar.array_tree(G, h), ar.array_tree(G, u), ar.elem_tree(G, pkey),
                             ar.array_tree(G, w), ar.array_tree(G, wp)),
                     8 * PRGHeuristic(params.prghash).min_no_seed_bytes())


def ccpos_prove(G, params: Params, g, h, u, pkey, w, wp, r, pi, s, rs):
    """hvzk/CCPoSW.java:75-150."""
    P = CCPoSBasicW(G, params.vbitlenro, params.ebitlenro, params.rbitlen, params.prghash)
    P.set_instance(g, h, u, pkey, w, wp, r, pi, s)
    seed = _ccpos_seed(G, params, g, h, u, pkey, w, wp)
    commitment = P.commit(seed, rs)
    cb = challenge(params.rohash, params.prefix(), bt.node(bt.leaf(seed), commitment), params.vbitlenro)
    reply = P.reply(int.from_bytes(cb, "big"))
    return commitment.to_bytes(), reply.to_bytes()


def ccpos_verify(G, params: Params, g, h, u, pkey, w, wp, commitment: bytes, reply: bytes, tv=None) -> bool:
    """hvzk/CCPoSW.java:160-260 with raisedu == null; mixnet/MixNetElGamalVerifyFiatShamirSession.java:757-841."""
    V = CCPoSBasicW(G, params.vbitlenro, params.ebitlenro, params.rbitlen, params.prghash)
    V.set_instance(g, h, u, pkey, w, wp)
    seed = _ccpos_seed(G, params, g, h, u, pkey, w, wp)
    V.set_batch_vector(seed)
    V.compute_AB()
    try:
        ctree = V.set_commitment(bt.read(commitment))
    except bt.EIOError:
        ctree = V.set_commitment(bt.leaf(b""))
    cb = challenge(params.rohash, params.prefix(), bt.node(bt.leaf(seed), ctree), params.vbitlenro)
    if tv is not None:
        tv["s"], tv["v"] = seed, int.from_bytes(cb, "big")
    V.set_challenge(int.from_bytes(cb, "big"))
    try:
        return V.verify(bt.read(reply))
    except bt.EIOError:
        return False


def precomp(G, params: Params, pkey, h, rs):
    """mixnet/PermutationCommitment.java:148-219 + :251-292 and mixnet/ShufflerElGamalSession.java:647-658:
    commit to a permutation of maxciph = len(h) generators, prove it, pre-compute re-encryption factors."""
    n = len(h)
    exponents = ar.ring_random_array(G, n, rs, params.rbitlen)
    idc = ar.g_mul(G, h, ar.g_exp(G, G.g, exponents))
    pi = ar.permutation_random(n, rs, params.rbitlen)
    u = ar.permute(idc, pi)
    posc_c, posc_r = posc_prove(G, params, G.g, h, u, exponents, pi, rs)
    shape = _ring_shape(pkey)        # exponents of the (possibly wide) ciphertext ring, component by component
    if isinstance(shape, tuple):
        s = tuple(ar.ring_random_array(G, n, rs, params.rbitlen) for _ in shape)
    else:
        s = ar.ring_random_array(G, n, rs, params.rbitlen)
    factors = ar.g_exp(G, pkey, s)
    state = {"exponents": exponents, "pi": pi, "u": u, "s": s, "factors": factors, "h": list(h)}
    return state, (ar.array_tree(G, u).to_bytes(), posc_c, posc_r)


def shrink(G, state: dict, n: int) -> bytes:
    """mixnet/PermutationCommitment.java:390-471 (l == j) + mixnet/ShufflerElGamalSession.java:673-760."""
    size = len(state["u"])
    keep = [False] * size
    for i in range(n):
        keep[state["pi"][i]] = True
    state["exponents"] = state["exponents"][:n]
    state["pi"] = perm_shrink(state["pi"], n)
    state["u"] = [x for x, k in zip(state["u"], keep) if k]
    state["h"] = state["h"][:n]
    state["s"] = ar.gmap(lambda col: col[:n], state["s"]) if isinstance(state["s"], tuple) else state["s"][:n]
    state["factors"] = ar.gmap(lambda col: col[:n], state["factors"])
    return bt.bool_array_leaf(keep).to_bytes()


def committed_shuffle(G, params: Params, pkey, state: dict, w, rs):
    """mixnet/ShufflerElGamalSession.java:771-822."""
    wp = ar.permute(ar.g_mul(G, w, state["factors"]), ar.perm_inv(state["pi"]))
    c, r = ccpos_prove(G, params, G.g, state["h"], state["u"], pkey, w, wp, state["exponents"], state["pi"],
                       state["s"], rs)
    return wp, {"output": ar.array_tree(G, wp).to_bytes(), "commitment": c, "reply": r}


# ---------------------------------------------------------------- a whole mix and its verification (vmnv)
# mixnet/MixNetElGamalSession (shuffling by the first `threshold` parties), elgamal/DistrElGamalSession.java:361-545
# (decryption), mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668 (verification); file names of the proof
# directory: mixnet/MixNetElGamalSession.java:381-446, mixnet/ShufflerElGamalSession.java:1077-1101,
# hvzk/PoSTW.java:281-307, elgamal/DistrElGamalSession.java:553-601, elgamal/DistrElGamal.java:245-255.
# Key generation is a Shamir sharing in the exponent dealt from `rs` (the DKG is control plane, out of scope);
# [VCR-mem] PolynomialInExponent.toByteTree() = node of its coefficient elements.
def _party_source(rs):
    prg = PRGHeuristic("sha256")
    seed = rs.get_bytes(prg.min_no_seed_bytes())
    return SeededRandomSource(seed)


def _dec_seed_data(G, g, w, coeffs, f, k):
    bt_in = bt.node(ar.elem_tree(G, g), ar.array_tree(G, w))
    pk_bt = bt.node([ar.elem_tree(G, c) for c in coeffs])
    df_bt = bt.node([ar.array_tree(G, f[l]) for l in range(1, k + 1)])
    return bt.node(bt_in, bt.node(pk_bt, df_bt))


def _eval_in_exponent(G, coeffs, l):
    acc, power = coeffs[0], 1
    for c in coeffs[1:]:
        power *= l
        acc = G.op_mul(acc, G.op_exp(c, power % G.q))
    return acc


def wide_key(pk, width: int):
    """elgamal/ProtocolElGamal.java:785-800: the basic key (g, y) as a key over width-omega ciphertexts."""
    return pk if width == 1 else ((pk[0],) * width, (pk[1],) * width)


def run_mix(G, params: Params, k: int, threshold: int, w, rs, auxsid=None, width: int = 1, mode: str = "mixing",
            maxciph=None):
    """Returns (proof directory as dict name -> bytes, plaintext elements -- the shuffled list for mode
    "shuffling").  `mode` is the type of the session (mixnet/MixNetElGamalSession.java:53-63): "mixing" (shuffle,
    then decrypt: :345-358), "shuffling" (:208-245) or "decryption" (:268-325).  `maxciph`: pre-computation for that
    many ciphertexts first (:161-187; ShufflerElGamalSession.java:534-672), so that the shuffle is the
    commitment-consistent one (:972-1033) over commitments shrunk to the actual number of ciphertexts."""
    if auxsid is not None:
        params = params.with_auxsid(auxsid)
    if mode not in ("mixing", "shuffling", "decryption"):
        raise ValueError(mode)
    q = G.q
    poly = [ar.ring_random_element(G, rs, params.rbitlen) for _ in range(threshold)]
    xs = {l: sum(a * pow(l, i, q) for i, a in enumerate(poly)) % q for l in range(1, k + 1)}
    coeffs = [G.op_exp(G.g, a) for a in poly]
    pk = (G.g, coeffs[0])
    wpk = wide_key(pk, width)
    d = {"version": params.version.encode(), "type": mode.encode(), "auxsid": params.auxsid.encode(),
         "width": str(width).encode(),
         "FullPublicKey.bt": ar.elem_tree(G, pk).to_bytes(),
         "proofs/PolynomialInExponent.bt": bt.node([ar.elem_tree(G, c) for c in coeffs]).to_bytes(),
         "Ciphertexts.bt": ar.array_tree(G, w).to_bytes(), "proofs/activethreshold": str(threshold).encode()}
    n = ar.size_of(w)
    inp = w
    if mode != "decryption" and maxciph is None:
        h = independent_generators(G, params.rohash, params.prefix(), "generators", n, params.rbitlen)
        for l in range(1, threshold + 1):
            out, proof = shuffle_and_prove(G, params, wpk, inp, h, _party_source(rs))
            d["ShuffledCiphertexts.bt" if l == threshold else "proofs/Ciphertexts%02d.bt" % l] = proof["output"]
            d["proofs/PermutationCommitment%02d.bt" % l] = proof["permutationCommitment"]
            d["proofs/PoSCommitment%02d.bt" % l] = proof["commitment"]
            d["proofs/PoSReply%02d.bt" % l] = proof["reply"]
            inp = out
    elif mode != "decryption":
        if maxciph < n:
            raise ValueError("more ciphertexts than pre-computed for")
        d["proofs/maxciph"] = str(maxciph).encode()
        h = independent_generators(G, params.rohash, params.prefix(), "generators", maxciph, params.rbitlen)
        states = {}
        for l in range(1, threshold + 1):       # pre-computation: every active party commits and proves (PoSC)
            states[l], (pc, c, r) = precomp(G, params, wpk, h, _party_source(rs))
            d["proofs/PermutationCommitment%02d.bt" % l] = pc
            d["proofs/PoSCCommitment%02d.bt" % l] = c
            d["proofs/PoSCReply%02d.bt" % l] = r
        for l in range(1, threshold + 1):       # shrink to the actual size, then the commitment-consistent shuffles
            d["proofs/KeepList%02d.bt" % l] = shrink(G, states[l], n)
        for l in range(1, threshold + 1):
            out, proof = committed_shuffle(G, params, wpk, states[l], inp, _party_source(rs))
            d["ShuffledCiphertexts.bt" if l == threshold else "proofs/Ciphertexts%02d.bt" % l] = proof["output"]
            d["proofs/CCPoSCommitment%02d.bt" % l] = proof["commitment"]
            d["proofs/CCPoSReply%02d.bt" % l] = proof["reply"]
            inp = out
    if mode == "shuffling":
        return d, inp
    if mode == "mixing":   # the output of the shuffle moves into the proofs (MixNetElGamalSession.java:294-303)
        d["proofs/Ciphertexts%02d.bt" % threshold] = d.pop("ShuffledCiphertexts.bt")
    # decryption
    u = inp[0]
    inv_factor = pow(prod_factor(q, k), -1, q)
    f = {l: decryption_factors(G, u, xs[l], inv_factor) for l in range(1, k + 1)}
    for l in f:
        d["proofs/DecryptionFactors%02d.bt" % l] = ar.array_tree(G, f[l]).to_bytes()
    correct = [False] + [True] * k
    combined = combine_decryption_factors(G, f, correct, k, threshold)
    prefix = params.prefix()
    seed = challenge(params.rohash, prefix, _dec_seed_data(G, G.g, inp, coeffs, f, k),
                     8 * PRGHeuristic(params.prghash).min_no_seed_bytes())
    ys = {l: G.op_exp(G.g, xs[l]) for l in xs}
    parties = {}
    for l in range(1, k + 1):
        E = DistrElGamalSessionBasic(G, l, k, threshold, params.ebitlenro, params.rbitlen, params.prghash, G.g, ys, u, xs[l])
        E.f = f
        E.set_batch_vector(seed)
        E.batch_input()
        d["proofs/DecrFactCommitment%02d.bt" % l] = E.commit(_party_source(rs)).to_bytes()
        parties[l] = E
    E1 = parties[1]
    for l in range(2, k + 1):
        E1.set_commitment(l, bt.from_bytes(d["proofs/DecrFactCommitment%02d.bt" % l]))
    cdata = bt.node(bt.leaf(seed), bt.node([E1.commitment_tree(l) for l in range(1, k + 1)]))
    v = int.from_bytes(challenge(params.rohash, prefix, cdata, params.vbitlenro), "big")
    for l in range(1, k + 1):
        d["proofs/DecrFactReply%02d.bt" % l] = parties[l].reply(v).to_bytes()
    d["proofs/CorrectIndices.bt"] = bt.leaf(bytes(1 if c else 0 for c in correct)).to_bytes()
    plain = ar.g_mul(G, inp[1], combined)
    d["Plaintexts.bt"] = ar.array_tree(G, plain).to_bytes()
    return d, plain


class MixVerificationError(Exception):
    pass


def _parse_int(raw: bytes) -> int:
    """Integer.parseInt: an optional sign and decimal digits, nothing else."""
    import re
    text = raw.decode("ascii", errors="replace")
    if re.fullmatch(r"[+-]?[0-9]{1,10}", text) is None:
        raise ValueError("not an integer")
    return int(text)


def verify_mix(G, params: Params, k: int, threshold: int, d: dict, expected_auxsid=None, expected_width=None,
               expected_type=None, dec: bool = True, posc: bool = True, ccpos: bool = True) -> dict:
    """The verdicts of mixnet/MixNetElGamalVerifyFiatShamirSession.verify (:1318-1668) for a proof of type "mixing",
    "shuffling" or "decryption", with or without pre-computation (`proofs/maxciph` present: proofs of shuffles of
    commitments, keep lists, commitment-consistent proofs of shuffles).  `dec`, `posc`, `ccpos`: what is verified
    (the -nodec / -noposc / -noccpos options of vmnv; mixnet/SessionParams.java).  A file that is malformed where the
    reference does not substitute trivial values is fail-stop."""
    try:
        return _verify_mix(G, params, k, threshold, d, expected_auxsid, expected_width, expected_type, dec, posc, ccpos)
    except (bt.EIOError, ar.FormatError, ValueError, IndexError, AttributeError, TypeError) as e:
        raise MixVerificationError("malformed proof directory: %s" % e)


def _first_array_size(G, t, width: int) -> int:
    """Number of elements of the ciphertext array serialised as `t` (readArray with size 0, :577-607)."""
    first = t.children[0]
    for _ in range((1 if width > 1 else 0) + (1 if hasattr(G, "coord_bytes") else 0)):
        first = first.children[0]
    if first.is_leaf():
        raise ar.FormatError("array expected")
    return len(first.children)


def _verify_mix(G, params: Params, k: int, threshold: int, d: dict, expected_auxsid=None, expected_width=None,
                expected_type=None, dec=True, posc=True, ccpos=True) -> dict:
    def need(name):
        if name not in d:
            raise MixVerificationError("missing " + name)
        return d[name]
    if need("version").decode() != params.version:
        raise MixVerificationError("version")
    # determineType (:329-358), determineSessionParams (:984-1005)
    typ = need("type").decode("ascii", errors="replace")
    if typ not in ("mixing", "shuffling", "decryption"):
        raise MixVerificationError("unknown type of proof")
    if expected_type is not None and typ != expected_type:
        raise MixVerificationError("type mismatch")
    # determineAuxsid (:369-395): read from the proof, validated, and part of the global prefix (:160); [VCR-mem]
    # Protocol.validateSid = letters, digits, underscores, spaces
    import re
    auxsid = need("auxsid").decode("ascii", errors="replace")
    if re.fullmatch(r"[A-Za-z0-9_ ]{1,1024}", auxsid) is None:
        raise MixVerificationError("auxsid")
    if expected_auxsid is not None and auxsid != expected_auxsid:
        raise MixVerificationError("auxsid mismatch")
    params = params.with_auxsid(auxsid)
    if typ == "shuffling":
        dec = False
    elif typ == "decryption":
        posc = ccpos = False
    width = 1
    if ccpos or dec:
        width = _parse_int(need("width"))
        if width < 1 or (expected_width is not None and width != expected_width):
            raise MixVerificationError("width")
    # the scalar test vectors of `vmnv -t` (mixnet/MixNetElGamalVerifyFiatShamirTool.java:82-224), in the order the
    # reference prints them: (name, party or None, value)
    vectors = []
    rep = {"type": typ, "shuffles": {}, "poscs": {}, "decryption": None, "vectors": vectors}

    def record(name, value, party=None):
        vectors.append((name, party, value.hex() if isinstance(value, (bytes, bytearray)) else str(value)))
    for name, value in (("par.k", k), ("par.lambda", threshold), ("par.n_e", params.ebitlenro), ("par.n_r", params.rbitlen),
                        ("par.n_v", params.vbitlenro), ("par.s_Gq", params.pgroup_string), ("par.version", params.version)):
        record(name, value)
    if ccpos or dec:
        record("par.omega", width)
    record("par.sid", params.sid)
    record("der.rho", params.prefix())
    try:
        pk = ar.parse_elem(G, bt.read(need("FullPublicKey.bt")), (None, None))
    except (ar.FormatError, bt.EIOError):
        raise MixVerificationError("keys")
    if pk[0] != G.g:
        raise MixVerificationError("basic public key is not the standard generator")
    coeffs, ys = None, None
    if dec:   # readMixServerPKeys :228-266
        try:
            t = bt.read(need("proofs/PolynomialInExponent.bt"))
            if t.is_leaf() or t.declared != threshold or len(t.children) != threshold:
                raise ar.FormatError("degree")
            coeffs = [ar.parse_elem(G, c) for c in t.children]
        except (ar.FormatError, bt.EIOError):
            raise MixVerificationError("keys")
        if pk[1] != coeffs[0]:
            raise MixVerificationError("mismatching keys")
        ys = {l: _eval_in_exponent(G, coeffs, l) for l in range(1, k + 1)}
    precomp_ = "proofs/maxciph" in d
    active = _parse_int(need("proofs/activethreshold"))
    if active > k or active < threshold:
        raise MixVerificationError("active threshold")
    record("par.lambda", active)
    basic_pk, pk = pk, wide_key(pk, width)
    # readCiphertexts :1017-1046
    w = None
    if ccpos or dec:
        if ccpos or typ == "decryption":
            name = "Ciphertexts.bt"
            raw = need(name)
        else:
            name = "proofs/Ciphertexts%02d.bt" % active
            raw = d.get(name)
        if raw is not None:
            try:
                ct = bt.read(raw)
                w = ar.parse_array(G, ct, _first_array_size(G, ct, width), pk)
            except (ar.FormatError, bt.EIOError, IndexError, AttributeError):
                raise MixVerificationError("ciphertexts")
            if ar.size_of(w) == 0:
                raise MixVerificationError("no ciphertexts")
    if posc or ccpos:
        # getMaxciph :541-548, deriveGenerators :556-576, getShrunkGenerators :1059-1068
        if precomp_:
            maxciph = _parse_int(need("proofs/maxciph"))
            # (no file of the directory could hold a commitment of that many elements)
            record("par.N_0", maxciph)
            if maxciph < 1 or maxciph > max(len(v) for v in d.values()) // (5 + (G.p.bit_length() + 7) // 8):
                raise MixVerificationError("maxciph")
        else:
            if w is None:
                raise MixVerificationError("no ciphertexts")
            maxciph = ar.size_of(w)
        h = independent_generators(G, params.rohash, params.prefix(), "generators", maxciph, params.rbitlen)
        shrunk = None
        if ccpos and precomp_:
            if ar.size_of(w) > maxciph:
                raise MixVerificationError("more ciphertexts than generators")
            shrunk = h[:ar.size_of(w)]
        inp, valid = w, 0
        for l in range(1, active + 1):
            verdict = True
            pcname = "proofs/PermutationCommitment%02d.bt" % l
            if (posc and precomp_ and not ccpos and pcname in d) or \
                    ("proofs/CCPoSCommitment%02d.bt" % l in d or "proofs/PoSCommitment%02d.bt" % l in d):
                try:   # readPermutationCommitment :626-641: fail-stop when missing or malformed
                    u = ar.parse_array(G, bt.read(need(pcname)), maxciph)
                except (ar.FormatError, bt.EIOError):
                    raise MixVerificationError("permutation commitment of party %d" % l)
                if posc and precomp_:   # verifyPoSC :652-705
                    tv = {}
                    ok = posc_verify(G, params, G.g, h, u, need("proofs/PoSCCommitment%02d.bt" % l),
                                     need("proofs/PoSCReply%02d.bt" % l), tv)
                    record("PoSC.s", tv["s"], l)
                    record("PoSC.v", tv["v"], l)
                    rep["poscs"][l] = ok
                    if not ok:
                        verdict = False
                        u = list(h)
                if ccpos:
                    n = ar.size_of(inp)
                    name = "proofs/Ciphertexts%02d.bt" % l
                    if l == active and name not in d:
                        name = "ShuffledCiphertexts.bt"
                    try:
                        out = ar.parse_array(G, bt.read(need(name)), n, pk)
                    except (ar.FormatError, bt.EIOError):
                        raise MixVerificationError("output of party %d" % l)
                    if precomp_:
                        # shrinkPermComm :714-745: a keep list that cannot be read or keeps the wrong number is fail-stop
                        kl = bt.read(need("proofs/KeepList%02d.bt" % l))
                        if not kl.is_leaf() or len(kl.value) != maxciph or any(x > 1 for x in kl.value):
                            raise MixVerificationError("keep list of party %d" % l)
                        if sum(kl.value) != n:
                            raise MixVerificationError("wrong number of true elements in keep list of party %d" % l)
                        su = [x for x, keep in zip(u, kl.value) if keep]
                        tv = {}
                        ok = ccpos_verify(G, params, G.g, shrunk, su, pk, inp, out,
                                          need("proofs/CCPoSCommitment%02d.bt" % l), need("proofs/CCPoSReply%02d.bt" % l), tv)
                        record("CCPoS.s", tv["s"], l)
                        record("CCPoS.v", tv["v"], l)
                        verdict = verdict and ok
                    else:
                        proof = {"output": d[name], "permutationCommitment": d[pcname],
                                 "commitment": need("proofs/PoSCommitment%02d.bt" % l),
                                 "reply": need("proofs/PoSReply%02d.bt" % l)}
                        tv = {}
                        verdict = verify_shuffle(G, params, pk, inp, h, proof, tv)
                        record("PoS.s", tv["s"], l)
                        record("PoS.v", tv["v"], l)
                    inp = out if verdict else inp
                rep["shuffles"][l] = verdict
                valid += 1 if verdict else 0
        rep["validProofs"] = valid
        rep["enoughValidProofs"] = valid >= threshold
        if dec:
            w = inp
    if not dec:
        rep["accepted"] = bool(rep.get("enoughValidProofs", True))
        return rep
    if w is None:
        raise MixVerificationError("no ciphertexts to decrypt")
    inp = w
    n = ar.size_of(inp)
    flags = bt.read(need("proofs/CorrectIndices.bt"))
    if not flags.is_leaf() or len(flags.value) != k + 1 or max(flags.value) > 1:
        raise MixVerificationError("correct indices")
    correct = [bool(x) for x in flags.value]
    if sum(correct[1:]) < threshold:
        raise MixVerificationError("too few correct decryption factors")
    u = inp[0]
    try:
        f = {l: ar.parse_array(G, bt.read(need("proofs/DecryptionFactors%02d.bt" % l)), n, pk[0] if width > 1 else None)
             for l in range(1, k + 1)}
    except (ar.FormatError, bt.EIOError):
        raise MixVerificationError("decryption factors")
    combined = combine_decryption_factors(G, f, correct, k, threshold)
    prefix = params.prefix()
    seed = challenge(params.rohash, prefix, _dec_seed_data(G, G.g, inp, coeffs, f, k),
                     8 * PRGHeuristic(params.prghash).min_no_seed_bytes())
    record("Dec.s", seed)
    V = DistrElGamalSessionBasic(G, 0, k, threshold, params.ebitlenro, params.rbitlen, params.prghash, G.g, ys, u)
    V.f = f
    V.set_batch_vector(seed)
    V.batch_input()
    V.batch_combined(combined)
    for l in range(1, k + 1):
        try:
            V.set_commitment(l, bt.read(need("proofs/DecrFactCommitment%02d.bt" % l)))
        except bt.EIOError:
            V.set_commitment(l, bt.leaf(b""))
    cdata = bt.node(bt.leaf(seed), bt.node([V.commitment_tree(l) for l in range(1, k + 1)]))
    v = int.from_bytes(challenge(params.rohash, prefix, cdata, params.vbitlenro), "big")
    record("Dec.v", v)
    for l in range(1, k + 1):
        try:
            V.set_reply(l, bt.read(need("proofs/DecrFactReply%02d.bt" % l)))
        except bt.EIOError:
            V.set_reply(l, bt.node([]))
    V.combine(correct)
    rep["decryption"] = V.verify_combined(basic_pk[1], v)
    if not rep["decryption"]:
        raise MixVerificationError("combined proof of decryption")
    computed = ar.g_mul(G, inp[1], combined)
    try:
        plain = ar.parse_array(G, bt.read(need("Plaintexts.bt")), n, pk[0] if width > 1 else None)
    except (ar.FormatError, bt.EIOError):
        raise MixVerificationError("plaintexts")
    rep["plaintexts"] = plain == computed
    if not rep["plaintexts"]:
        raise MixVerificationError("plaintexts are incorrect")
    rep["accepted"] = bool(rep.get("enoughValidProofs", True))
    return rep
