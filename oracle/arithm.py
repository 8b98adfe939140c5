"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement, on Python integers, of the `com.verificatum.arithm` semantics the hot path
relies on.  The classes live in the un-vendored jar verificatum-vcr 3.1.0 (/root/reference
configure.ac:35); what is restated here are the rules of SURVEY.md §8c ([VCR-mem]) plus plain
modular arithmetic, each anchored on the reference call site that uses it:

  group arrays   exp / expProd / mul / prod / permute / shiftPush      hvzk/PoSBasicTW.java:407-482,546-700,1000-1066
  ring arrays    recLin / prods / mulAdd / innerProduct / sum / prod   hvzk/PoSBasicTW.java:583-604,861-878
  random arrays  pRing.randomElementArray, LargeIntegerArray.random,   hvzk/PoSBasicTW.java:446,470-474,533-538
                 pGroup.randomElementArray, Permutation.random         distr/IndependentGeneratorsRO.java:129,
                                                                       mixnet/ShufflerElGamalSession.java:400-409
  encodings      fixed-width two's-complement leaves, column-wise      elgamal/ProtocolElGamalInterfaceRaw.java:53-56
                 product arrays

PARITY UNPINNED against the Java/GMP path itself: the reference tree holds no golden vector
for any of these; what pins this file is (i) exact integer arithmetic having one
right answer, (ii) the byte-tree / PRG / RO known-answer values in tests/golden/, (iii) the
in-tree ModPGroup fixture, (iv) prove/verify self-consistency.

Structures: a group element is an int; a product element is a tuple of elements; an array is
a list of ints; a product array is a tuple of arrays (column-wise).  Ring elements likewise.
"""
from __future__ import annotations

from typing import List, Sequence

from . import bytetree as bt


class ModPGroup:
    """Subgroup of order q of Z_p^*.  A group exposes one, op_mul / op_inv / op_exp on single elements
    (ints here, ECPoint for oracle.ec.ECqPGroup) and its own element / array encodings, so that the
    array functions and the protocol restatements below are written once for both."""

    def __init__(self, p: int, q: int, g: int):
        assert pow(g, q, p) == 1 and g != 1
        self.p, self.q, self.g = p, q, g
        self.elem_bytes = p.bit_length() // 8 + 1   # BigInteger.toByteArray() length of p
        self.ring_bytes = q.bit_length() // 8 + 1
        self.cofactor = (p - 1) // q
        self.one = 1

    def contains(self, x: int) -> bool:
        return 0 < x < self.p and pow(x, self.q, self.p) == 1

    def op_mul(self, a: int, b: int) -> int:
        return a * b % self.p

    def op_inv(self, a: int) -> int:
        return pow(a, -1, self.p)

    def op_exp(self, a: int, e: int) -> int:
        return pow(a, e, self.p)

    # encodings: a leaf of elem_bytes bytes per element, an array is a node of such leaves
    def leaf_tree(self, x: int) -> bt.ByteTree:
        return bt.int_leaf(x, self.elem_bytes)

    def leaf_array_tree(self, arr) -> bt.ByteTree:
        return bt.node([bt.int_leaf(x, self.elem_bytes) for x in arr])

    def parse_leaf(self, t: bt.ByteTree) -> int:
        if not t.is_leaf() or len(t.value) != self.elem_bytes:
            raise FormatError("element length")
        x = bt.bytes_to_int(t.value)
        if not self.contains(x):
            raise FormatError("not a group element")
        return x

    def parse_leaf_array(self, t: bt.ByteTree, size: int):
        if t.is_leaf() or t.declared != size or len(t.children) != size:
            raise FormatError("array size")
        return [self.parse_leaf(c) for c in t.children]

    def random_array(self, n: int, rs, stat_dist: int):
        """pGroup.randomElementArray(n, prg, statDist) (distr/IndependentGeneratorsRO.java:129):
        t^((p-1)/q) for wide random t."""
        bits = self.p.bit_length() + stat_dist
        w = (bits + 7) // 8
        return [pow(_masked_int(rs.get_bytes(w), bits) % self.p, self.cofactor, self.p) for _ in range(n)]


# ---------------------------------------------------------------- structure helpers
def is_product(x) -> bool:
    return isinstance(x, tuple)


def gmap(f, x, *others):
    """Apply f to the leaves (ints or lists) of equally shaped nested tuples."""
    if isinstance(x, tuple):
        return tuple(gmap(f, c, *[o[i] if isinstance(o, tuple) else o for o in others]) for i, c in enumerate(x))
    return f(x, *others)


def leaves(x) -> list:
    if isinstance(x, tuple):
        out = []
        for c in x:
            out += leaves(c)
        return out
    return [x]


def size_of(arr) -> int:
    return len(leaves(arr)[0])


# ---------------------------------------------------------------- byte trees
def elem_tree(G: ModPGroup, x) -> bt.ByteTree:
    if isinstance(x, tuple):
        return bt.node([elem_tree(G, c) for c in x])
    return G.leaf_tree(x)


def array_tree(G: ModPGroup, arr) -> bt.ByteTree:
    if isinstance(arr, tuple):
        return bt.node([array_tree(G, c) for c in arr])
    return G.leaf_array_tree(arr)


def ring_tree(G: ModPGroup, x) -> bt.ByteTree:
    if isinstance(x, tuple):
        return bt.node([ring_tree(G, c) for c in x])
    return bt.int_leaf(x, G.ring_bytes)


def ring_array_tree(G: ModPGroup, arr) -> bt.ByteTree:
    if isinstance(arr, tuple):
        return bt.node([ring_array_tree(G, c) for c in arr])
    return bt.node([bt.int_leaf(x, G.ring_bytes) for x in arr])


class FormatError(ValueError):
    """ArithmFormatException / EIOException."""


def parse_elem(G: ModPGroup, t: bt.ByteTree, shape=None) -> int:
    if isinstance(shape, tuple):
        if t.is_leaf() or t.declared != len(shape) or len(t.children) != len(shape):
            raise FormatError("arity")
        return tuple(parse_elem(G, c, s) for c, s in zip(t.children, shape))
    return G.parse_leaf(t)


def parse_array(G: ModPGroup, t: bt.ByteTree, size: int, shape=None):
    if isinstance(shape, tuple):
        if t.is_leaf() or t.declared != len(shape) or len(t.children) != len(shape):
            raise FormatError("arity")
        return tuple(parse_array(G, c, size, s) for c, s in zip(t.children, shape))
    return G.parse_leaf_array(t, size)


def parse_ring(G: ModPGroup, t: bt.ByteTree, shape=None):
    if isinstance(shape, tuple):
        if t.is_leaf() or t.declared != len(shape) or len(t.children) != len(shape):
            raise FormatError("arity")
        return tuple(parse_ring(G, c, s) for c, s in zip(t.children, shape))
    if not t.is_leaf() or len(t.value) != G.ring_bytes:
        raise FormatError("ring element length")
    x = bt.bytes_to_int(t.value)
    if not 0 <= x < G.q:
        raise FormatError("ring element out of range")
    return x


def parse_ring_array(G: ModPGroup, t: bt.ByteTree, size: int):
    if t.is_leaf() or t.declared != size or len(t.children) != size:
        raise FormatError("array size")
    return [parse_ring(G, c) for c in t.children]


# ---------------------------------------------------------------- random objects
def _masked_int(raw: bytes, bits: int) -> int:
    return int.from_bytes(raw, "big") & ((1 << bits) - 1)


def ring_random_element(G: ModPGroup, rs, stat_dist: int) -> int:
    """pRing.randomElement(rs, statDist): ceil((|q|+statDist)/8) bytes, masked, mod q."""
    bits = G.q.bit_length() + stat_dist
    return _masked_int(rs.get_bytes((bits + 7) // 8), bits) % G.q


def ring_random_array(G: ModPGroup, n: int, rs, stat_dist: int) -> List[int]:
    """pRing.randomElementArray(n, rs, statDist) (hvzk/PoSBasicTW.java:446,571,621)."""
    return [ring_random_element(G, rs, stat_dist) for _ in range(n)]


def lia_random(n: int, bits: int, rs) -> List[int]:
    """LargeIntegerArray.random(n, bits, rs) (hvzk/PoSBasicTW.java:472-474,535-536)."""
    w = (bits + 7) // 8
    return [_masked_int(rs.get_bytes(w), bits) for _ in range(n)]


def group_random_array(G: ModPGroup, n: int, rs, stat_dist: int):
    return G.random_array(n, rs, stat_dist)


def permutation_random(n: int, rs, stat_dist: int) -> List[int]:
    """Permutation.random(n, rs, statDist) (mixnet/ShufflerElGamalSession.java:408-409): n keys of
    ceil(log2 n)+statDist bits; the table sends the i-th smallest key's index to i."""
    bits = max(1, (n - 1).bit_length()) + stat_dist
    w = (bits + 7) // 8
    keys = [_masked_int(rs.get_bytes(w), bits) for _ in range(n)]
    order = sorted(range(n), key=lambda i: (keys[i], i))
    table = [0] * n
    for rank, i in enumerate(order):
        table[i] = rank
    return table


def perm_inv(table: Sequence[int]) -> List[int]:
    inv = [0] * len(table)
    for i, t in enumerate(table):
        inv[t] = i
    return inv


def permute(arr, table):
    """array.permute(pi): result[pi(i)] = this[i]."""
    def one(col):
        out = [None] * len(col)
        for i, x in enumerate(col):
            out[table[i]] = x
        return out
    return gmap(one, arr)


# ---------------------------------------------------------------- group array operations
def g_mul(G, a, b):
    return gmap(lambda x, y: [G.op_mul(u, v) for u, v in zip(x, y)] if isinstance(x, list) else G.op_mul(x, y), a, b)


def g_inv(G, a):
    return gmap(lambda x: [G.op_inv(u) for u in x] if isinstance(x, list) else G.op_inv(x), a)


def g_exp(G, base, e):
    """base: element (possibly product) or array; e: ring element, ring array, or product thereof.
    An exponent shaped like the base acts component-wise, otherwise on every component."""
    if isinstance(base, tuple):
        if isinstance(e, tuple) and _same_shape(base, e):
            return tuple(g_exp(G, b, x) for b, x in zip(base, e))
        return tuple(g_exp(G, b, e) for b in base)
    if isinstance(base, list):
        if isinstance(e, list):
            return [G.op_exp(b, x) for b, x in zip(base, e)]
        return [G.op_exp(b, e) for b in base]
    if isinstance(e, list):
        return [G.op_exp(base, x) for x in e]
    return G.op_exp(base, e)


def _same_shape(a, b) -> bool:
    if isinstance(a, tuple) != isinstance(b, tuple):
        return False
    if isinstance(a, tuple):
        return len(a) == len(b) and all(_same_shape(x, y) for x, y in zip(a, b))
    return True


def g_exp_prod(G, arr, e: List[int]):
    """array.expProd(e) = prod_i arr[i]^e[i] (component-wise for product arrays)."""
    def one(col):
        acc = G.one
        for x, k in zip(col, e):
            acc = G.op_mul(acc, G.op_exp(x, k))
        return acc
    return gmap(one, arr)


def g_prod(G, arr):
    def one(col):
        acc = G.one
        for x in col:
            acc = G.op_mul(acc, x)
        return acc
    return gmap(one, arr)


def shift_push(arr, first):
    return gmap(lambda col, f: [f] + col[:-1], arr, first)


# ---------------------------------------------------------------- ring array operations (Z_q)
def r_rec_lin(G, b: List[int], e: List[int]):
    """b.recLin(e): x[0] = b[0]; x[i] = x[i-1]*e[i] + b[i] (hvzk/PoSBasicTW.java:583-598)."""
    x = [b[0] % G.q]
    for i in range(1, len(b)):
        x.append((x[-1] * e[i] + b[i]) % G.q)
    return x, x[-1]


def r_prods(G, e: List[int]) -> List[int]:
    out, acc = [], 1
    for v in e:
        acc = acc * v % G.q
        out.append(acc)
    return out


def r_inner(G, a, b) -> int:
    if isinstance(a, tuple):
        return tuple(r_inner(G, c, b) for c in a)
    return sum(x * y for x, y in zip(a, b)) % G.q
