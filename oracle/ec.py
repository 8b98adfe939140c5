"""ORACLE (test infrastructure only -- never imported by the product path).

Elliptic-curve groups of prime order over prime fields, y^2 = x^3 + a x + b: the CPU restatement
of `com.verificatum.arithm.ECqPGroup` (un-vendored jar verificatum-vcr 3.1.0, /root/reference
configure.ac:35; its natives verificatum-vec / vecj are named at demo/mixnet/.conf:143-145) on
Python integers.  The reference's own default group is a NIST curve: demo/mixnet/.checkbaseconf
(P-224), demo/mixnet/benchmarks/bench_config:33-50 (P-256); BASELINE.json config 5 is P-256.

PARITY UNPINNED against the Java path: the reference tree holds no EC fixture at all.  What pins
this file: the curve constants are checked against their defining equations (generator on the curve,
n * G = O), exact integer arithmetic, and the group laws in tests/test_oracle_ec.py.  The
ENCODINGS below are [VCR-mem]:
    element        node(leaf(x), leaf(y)), each coordinate a fixed-width two's-complement leaf of
                   bytelen(p) bytes; the unit element is (-1, -1)
    array          node(node(x_0 .. x_{n-1}), node(y_0 .. y_{n-1}))
    randomElementArray   per element: draw ceil((|p| + statDist)/8) bytes, reduce mod p to x, accept
                   if x^3 + ax + b is a square, take the smaller root y; else draw again
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

from . import bytetree as bt
from .arithm import FormatError, _masked_int


@dataclass(frozen=True)
class ECPoint:
    x: Optional[int]
    y: Optional[int]

    def is_unit(self) -> bool:
        return self.x is None


UNIT = ECPoint(None, None)

# FIPS 186-4 D.1.2.3
P256 = dict(
    name="P-256",
    p=0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF,
    a=0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFC,
    b=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
    gx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
    gy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
    n=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
)

# SEC 2, section 2.4.1 (a = 0: exercises the general doubling formula and a modulus without the NIST shape)
SECP256K1 = dict(
    name="secp256k1",
    p=2 ** 256 - 2 ** 32 - 977, a=0, b=7,
    gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
)
CURVES = {"P-256": P256, "secp256k1": SECP256K1}


class ECqPGroup:
    """Same surface as oracle.arithm.ModPGroup (one, op_mul/op_inv/op_exp, encodings)."""

    def __init__(self, name="P-256", p=None, a=None, b=None, gx=None, gy=None, n=None):
        if p is None:
            c = CURVES[name]
            p, a, b, gx, gy, n = c["p"], c["a"], c["b"], c["gx"], c["gy"], c["n"]
        self.name, self.p, self.a, self.b, self.q = name, p, a % p, b % p, n
        self.g = ECPoint(gx, gy)
        self.one = UNIT
        self.coord_bytes = p.bit_length() // 8 + 1
        self.ring_bytes = n.bit_length() // 8 + 1
        assert self.on_curve(self.g) and self.op_exp(self.g, n).is_unit()

    # ---- group law (Jacobian internally, affine results)
    def on_curve(self, P: ECPoint) -> bool:
        if P.is_unit():
            return True
        return 0 <= P.x < self.p and 0 <= P.y < self.p and (P.y * P.y - (P.x ** 3 + self.a * P.x + self.b)) % self.p == 0

    def contains(self, P: ECPoint) -> bool:
        return self.on_curve(P)  # prime order: every curve point is in the group

    def _dbl(self, X, Y, Z):
        p = self.p
        if Z == 0 or Y == 0:
            return 0, 1, 0
        S = 4 * X * Y * Y % p
        M = (3 * X * X + self.a * pow(Z, 4, p)) % p
        X3 = (M * M - 2 * S) % p
        Y3 = (M * (S - X3) - 8 * pow(Y, 4, p)) % p
        Z3 = 2 * Y * Z % p
        return X3, Y3, Z3

    def _add(self, X1, Y1, Z1, X2, Y2, Z2):
        p = self.p
        if Z1 == 0:
            return X2, Y2, Z2
        if Z2 == 0:
            return X1, Y1, Z1
        Z1Z1, Z2Z2 = Z1 * Z1 % p, Z2 * Z2 % p
        U1, U2 = X1 * Z2Z2 % p, X2 * Z1Z1 % p
        S1, S2 = Y1 * Z2 * Z2Z2 % p, Y2 * Z1 * Z1Z1 % p
        if U1 == U2:
            if S1 != S2:
                return 0, 1, 0
            return self._dbl(X1, Y1, Z1)
        H, R = (U2 - U1) % p, (S2 - S1) % p
        HH = H * H % p
        HHH = H * HH % p
        V = U1 * HH % p
        X3 = (R * R - HHH - 2 * V) % p
        Y3 = (R * (V - X3) - S1 * HHH) % p
        Z3 = H * Z1 * Z2 % p
        return X3, Y3, Z3

    def _affine(self, X, Y, Z) -> ECPoint:
        if Z == 0:
            return UNIT
        zi = pow(Z, -1, self.p)
        z2 = zi * zi % self.p
        return ECPoint(X * z2 % self.p, Y * z2 * zi % self.p)

    @staticmethod
    def _jac(P: ECPoint):
        return (0, 1, 0) if P.is_unit() else (P.x, P.y, 1)

    def op_mul(self, A: ECPoint, B: ECPoint) -> ECPoint:
        return self._affine(*self._add(*self._jac(A), *self._jac(B)))

    def op_inv(self, A: ECPoint) -> ECPoint:
        return A if A.is_unit() else ECPoint(A.x, (-A.y) % self.p)

    def op_exp(self, A: ECPoint, e: int) -> ECPoint:
        e %= self.q if hasattr(self, "q") and self.q else e + 1
        acc = (0, 1, 0)
        base = self._jac(A)
        for bit in bin(e)[2:] if e else "":
            acc = self._dbl(*acc)
            if bit == "1":
                acc = self._add(*acc, *base)
        return self._affine(*acc)

    # ---- encodings ([VCR-mem])
    def _coord(self, v: Optional[int]) -> bytes:
        return bt.int_to_bytes(-1 if v is None else v, self.coord_bytes)

    def leaf_tree(self, P: ECPoint) -> bt.ByteTree:
        return bt.node(bt.leaf(self._coord(P.x)), bt.leaf(self._coord(P.y)))

    def leaf_array_tree(self, arr) -> bt.ByteTree:
        return bt.node(bt.node([bt.leaf(self._coord(P.x)) for P in arr]), bt.node([bt.leaf(self._coord(P.y)) for P in arr]))

    def _decode(self, xb: bytes, yb: bytes) -> ECPoint:
        if len(xb) != self.coord_bytes or len(yb) != self.coord_bytes:
            raise FormatError("coordinate length")
        x, y = bt.bytes_to_int(xb), bt.bytes_to_int(yb)
        if x == -1 and y == -1:
            return UNIT
        P = ECPoint(x, y)
        if not (0 <= x < self.p and 0 <= y < self.p and self.on_curve(P)):
            raise FormatError("not a point of the curve")
        return P

    def parse_leaf(self, t: bt.ByteTree) -> ECPoint:
        if t.is_leaf() or t.declared != 2 or len(t.children) != 2 or not t.children[0].is_leaf() or not t.children[1].is_leaf():
            raise FormatError("point arity")
        return self._decode(t.children[0].value, t.children[1].value)

    def parse_leaf_array(self, t: bt.ByteTree, size: int):
        if t.is_leaf() or len(t.children) != 2:
            raise FormatError("point array arity")
        xs, ys = t.children
        if xs.is_leaf() or ys.is_leaf() or len(xs.children) != size or len(ys.children) != size:
            raise FormatError("array size")
        out = []
        for cx, cy in zip(xs.children, ys.children):
            if not cx.is_leaf() or not cy.is_leaf():
                raise FormatError("coordinate is not a leaf")
            out.append(self._decode(cx.value, cy.value))
        return out

    # ---- random elements ([VCR-mem]; distr/IndependentGeneratorsRO.java:129 calls this for the generators)
    def sqrt(self, v: int) -> Optional[int]:
        p = self.p
        assert p % 4 == 3
        r = pow(v, (p + 1) // 4, p)
        return r if r * r % p == v % p else None

    def random_array(self, n: int, rs, stat_dist: int) -> List[ECPoint]:
        bits = self.p.bit_length() + stat_dist
        w = (bits + 7) // 8
        out = []
        while len(out) < n:
            x = _masked_int(rs.get_bytes(w), bits) % self.p
            y = self.sqrt((x * x * x + self.a * x + self.b) % self.p)
            if y is None:
                continue
            out.append(ECPoint(x, min(y, self.p - y)))
        return out
