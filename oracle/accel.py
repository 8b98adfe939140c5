"""ORACLE (test infrastructure).  Optional GMP acceleration of the oracle's array operations
through oracle/libvmxref.so (cpu_ref.c): same results as the Python-integer versions in
oracle/arithm.py (checked in tests/test_oracle_accel.py), ~10x faster per exponentiation and
multi-threaded, which lets parity tests and the CPU baseline run at thousands of elements."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libvmxref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libvmxref.so not built (make -C oracle)")
        L = C.CDLL(path)
        vp, sz, ci = C.c_char_p, C.c_size_t, C.c_int
        L.ref_powm_array.argtypes = [vp, vp, ci, vp, ci, sz, vp, sz, sz, ci]
        L.ref_mul_array.argtypes = [vp, vp, vp, sz, vp, sz, ci]
        L.ref_powm_array_neg.argtypes = [vp, vp, vp, sz, vp, sz, sz, ci]
        L.ref_powm_array_neg.restype = None
        L.ref_fixed_table_create.argtypes = [vp, vp, sz, ci, ci]
        L.ref_fixed_table_create.restype = C.c_void_p
        L.ref_fixed_table_free.argtypes = [C.c_void_p]
        L.ref_fixed_table_free.restype = None
        L.ref_fixed_exp.argtypes = [vp, C.c_void_p, vp, sz, vp, sz, sz, ci]
        L.ref_jacobi_array.argtypes = [vp, vp, sz, vp, sz]
        L.ref_jacobi_array.restype = None
        L.ref_expprod.argtypes = [vp, vp, vp, sz, vp, sz, sz, ci, ci]
        for f in (L.ref_powm_array, L.ref_mul_array, L.ref_fixed_exp, L.ref_expprod):
            f.restype = None
        _LIB = L
    return _LIB


def cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _pack(vals, w):
    return b"".join(v.to_bytes(w, "big") for v in vals)


def _unpack(buf, n, w):
    return [int.from_bytes(buf[i * w:(i + 1) * w], "big") for i in range(n)]


class Accel:
    """GMP-backed versions of oracle.arithm's g_exp / g_exp_prod / g_mul for basic (non-product) operands."""

    def __init__(self, p: int, threads: int | None = None, fixed_window: int = 8, spowm_width: int = 7, q: int = 0):
        self.p = p
        self.q = q
        self.ew = (p.bit_length() + 7) // 8
        self.mod = p.to_bytes(self.ew, "big")
        self.threads = threads or cores()
        self.w = fixed_window
        self.k = spowm_width
        self.tables = {}

    def __del__(self):
        try:
            for h in self.tables.values():
                lib().ref_fixed_table_free(h)
        except Exception:
            pass

    def members(self, vals):
        n = len(vals)
        out = C.create_string_buffer(n)
        lib().ref_jacobi_array(out, _pack(vals, self.ew), n, self.mod, self.ew)
        return [b == 1 for b in out.raw]

    def exp_fixed(self, base: int, exps, ebits: int):
        n = len(exps)
        key = (base, ebits)
        if key not in self.tables:  # VCR keeps one fpowm table per fixed base as well
            self.tables[key] = lib().ref_fixed_table_create(base.to_bytes(self.ew, "big"), self.mod, self.ew, ebits, self.w)
        out = C.create_string_buffer(n * self.ew)
        lib().ref_fixed_exp(out, self.tables[key], _pack(exps, self.ew), n, self.mod, self.ew, self.ew, self.threads)
        return _unpack(out.raw, n, self.ew)

    def exp_var(self, bases, exps):
        n = len(bases)
        e_scalar = not isinstance(exps, list)
        out = C.create_string_buffer(n * self.ew)
        if e_scalar and self.q and 0 < self.q - exps < 1 << 64:   # a small negative integer mod q
            lib().ref_powm_array_neg(out, _pack(bases, self.ew), (self.q - exps).to_bytes(self.ew, "big"), n, self.mod,
                                     self.ew, self.ew, self.threads)
            return _unpack(out.raw, n, self.ew)
        lib().ref_powm_array(out, _pack(bases, self.ew), 0, _pack([exps] if e_scalar else exps, self.ew),
                             1 if e_scalar else 0, n, self.mod, self.ew, self.ew, self.threads)
        return _unpack(out.raw, n, self.ew)

    def expprod(self, bases, exps) -> int:
        n = len(bases)
        out = C.create_string_buffer(self.ew)
        lib().ref_expprod(out, _pack(bases, self.ew), _pack(exps, self.ew), n, self.mod, self.ew, self.ew, self.k,
                          self.threads)
        return int.from_bytes(out.raw, "big")

    def mul(self, a, b):
        n = len(a)
        out = C.create_string_buffer(n * self.ew)
        lib().ref_mul_array(out, _pack(a, self.ew), _pack(b, self.ew), n, self.mod, self.ew, self.threads)
        return _unpack(out.raw, n, self.ew)


def install(G, threads: int | None = None):
    """Route oracle.arithm's heavy array operations for group G through GMP.  Returns an undo()."""
    from . import arithm as ar
    acc = Accel(G.p, threads, q=G.q)
    orig = (ar.g_exp, ar.g_exp_prod, ar.g_mul)

    def g_exp(GG, base, e):
        if GG is not G or isinstance(base, tuple):
            if isinstance(base, tuple):
                if isinstance(e, tuple) and ar._same_shape(base, e):
                    return tuple(g_exp(GG, b, x) for b, x in zip(base, e))
                return tuple(g_exp(GG, b, e) for b in base)
            return orig[0](GG, base, e)
        if isinstance(base, list):
            return acc.exp_var(base, e) if base else []
        if isinstance(e, list):
            return acc.exp_fixed(base, e, GG.q.bit_length()) if e else []
        return acc.exp_var([base], [e])[0]

    def g_exp_prod(GG, arr, e):
        if GG is not G:
            return orig[1](GG, arr, e)
        return ar.gmap(lambda col: acc.expprod(col, e) if col else 1, arr)

    def g_mul(GG, a, b):
        if GG is not G:
            return orig[2](GG, a, b)
        return ar.gmap(lambda x, y: acc.mul(x, y) if isinstance(x, list) else x * y % GG.p, a, b)

    ar.g_exp, ar.g_exp_prod, ar.g_mul = g_exp, g_exp_prod, g_mul
    orig_parse = ar.parse_array

    def parse_array(GG, t, size, shape=None):
        """Same acceptance set as oracle.arithm.parse_array; for a safe prime the subgroup test
        x^q == 1 is evaluated as Legendre symbol == 1 (equivalent for p = 2q + 1)."""
        if GG is not G or isinstance(shape, tuple) or GG.p != 2 * GG.q + 1:
            if isinstance(shape, tuple) and GG is G:
                if t.is_leaf() or len(t.children) != len(shape):
                    raise ar.FormatError("arity")
                return tuple(parse_array(GG, c, size, s) for c, s in zip(t.children, shape))
            return orig_parse(GG, t, size, shape)
        if t.is_leaf() or len(t.children) != size:
            raise ar.FormatError("array size")
        vals = []
        for c in t.children:
            if not c.is_leaf() or len(c.value) != GG.elem_bytes:
                raise ar.FormatError("element length")
            v = int.from_bytes(c.value, "big", signed=True)
            if not 0 < v < GG.p:
                raise ar.FormatError("not a group element")
            vals.append(v)
        if not all(acc.members(vals)):
            raise ar.FormatError("not a group element")
        return vals

    ar.parse_array = parse_array
    # single elements (A', C', D', F' of a commitment): the same Legendre-symbol test instead of a Python pow(x, q, p)
    # per element (0.11 s each at 3072 bits -- a constant 0.6 s per step of the CPU arm)
    had_contains = "contains" in G.__dict__
    if G.p == 2 * G.q + 1:
        G.contains = lambda x: 0 < x < G.p and acc.members([x])[0]

    def undo():
        ar.g_exp, ar.g_exp_prod, ar.g_mul = orig
        ar.parse_array = orig_parse
        if not had_contains:
            G.__dict__.pop("contains", None)
    return undo


# ---------------------------------------------------------------------------------------------- curve groups
def _lib_ec():
    L = lib()
    if not getattr(L, "_ec_ready", False):
        vp, sz, ci = C.c_char_p, C.c_size_t, C.c_int
        L.ref_ec_exp_array.argtypes = [vp, vp, ci, vp, ci, sz, vp, sz, ci]
        L.ref_ec_mul_array.argtypes = [vp, vp, vp, sz, vp, ci]
        L.ref_ec_fixed_table_create.argtypes = [vp, vp, ci, ci]
        L.ref_ec_fixed_table_create.restype = C.c_void_p
        L.ref_ec_fixed_table_free.argtypes = [C.c_void_p]
        L.ref_ec_fixed_table_free.restype = None
        L.ref_ec_fixed_exp.argtypes = [vp, C.c_void_p, vp, sz, vp, sz, ci]
        L.ref_ec_expprod.argtypes = [vp, vp, vp, sz, vp, sz, ci, ci]
        for f in (L.ref_ec_exp_array, L.ref_ec_mul_array, L.ref_ec_fixed_exp, L.ref_ec_expprod):
            f.restype = None
        L._ec_ready = True
    return L


class AccelEC:
    """GMP-backed versions of the oracle's array operations over an oracle.ec.ECqPGroup (cpu_ref_ec.c): the CPU
    baseline of the curve workloads.  Points cross as x || y (32 bytes each), the unit element as 64 bytes 0xff."""

    XW = 32

    def __init__(self, G, threads: int | None = None, fixed_window: int = 8, smul_width: int = 5):
        from .ec import ECPoint, UNIT
        self.G, self.ECPoint, self.UNIT = G, ECPoint, UNIT
        self.curve = G.p.to_bytes(32, "big") + (G.a % G.p).to_bytes(32, "big") + (G.b % G.p).to_bytes(32, "big")
        self.threads = threads or cores()
        self.w, self.k = fixed_window, smul_width
        self.tables = {}

    def __del__(self):
        try:
            for h in self.tables.values():
                _lib_ec().ref_ec_fixed_table_free(h)
        except Exception:
            pass

    def _pt(self, P) -> bytes:
        return b"\xff" * 64 if P.is_unit() else P.x.to_bytes(32, "big") + P.y.to_bytes(32, "big")

    def _pack(self, pts) -> bytes:
        return b"".join(self._pt(P) for P in pts)

    def _unpack(self, raw: bytes, n: int):
        out = []
        for i in range(n):
            c = raw[64 * i:64 * i + 64]
            out.append(self.UNIT if c == b"\xff" * 64 else
                       self.ECPoint(int.from_bytes(c[:32], "big"), int.from_bytes(c[32:], "big")))
        return out

    def _scalars(self, exps) -> bytes:
        q = self.G.q
        return b"".join((e % q).to_bytes(self.XW, "big") for e in exps)

    def exp_fixed(self, base, exps):
        n = len(exps)
        key = self._pt(base)
        if key not in self.tables:
            self.tables[key] = _lib_ec().ref_ec_fixed_table_create(key, self.curve, self.G.q.bit_length(), self.w)
        out = C.create_string_buffer(64 * n)
        _lib_ec().ref_ec_fixed_exp(out, self.tables[key], self._scalars(exps), n, self.curve, self.XW, self.threads)
        return self._unpack(out.raw, n)

    def exp_var(self, bases, exps):
        n = len(bases)
        e_scalar = not isinstance(exps, list)
        out = C.create_string_buffer(64 * n)
        _lib_ec().ref_ec_exp_array(out, self._pack(bases), 0, self._scalars([exps] if e_scalar else exps),
                                   1 if e_scalar else 0, n, self.curve, self.XW, self.threads)
        return self._unpack(out.raw, n)

    def expprod(self, bases, exps):
        out = C.create_string_buffer(64)
        _lib_ec().ref_ec_expprod(out, self._pack(bases), self._scalars(exps), len(bases), self.curve, self.XW, self.k,
                                 self.threads)
        return self._unpack(out.raw, 1)[0]

    def mul(self, a, b):
        n = len(a)
        out = C.create_string_buffer(64 * n)
        _lib_ec().ref_ec_mul_array(out, self._pack(a), self._pack(b), n, self.curve, self.threads)
        return self._unpack(out.raw, n)


def install_ec(G, threads: int | None = None):
    """Route oracle.arithm's heavy array operations for the curve group G through cpu_ref_ec.c.  Returns an undo()."""
    from . import arithm as ar
    acc = AccelEC(G, threads)
    orig = (ar.g_exp, ar.g_exp_prod, ar.g_mul)

    def g_exp(GG, base, e):
        if GG is not G:
            return orig[0](GG, base, e)
        if isinstance(base, tuple):
            if isinstance(e, tuple) and ar._same_shape(base, e):
                return tuple(g_exp(GG, b, x) for b, x in zip(base, e))
            return tuple(g_exp(GG, b, e) for b in base)
        if isinstance(base, list):
            return acc.exp_var(base, e) if base else []
        if isinstance(e, list):
            return acc.exp_fixed(base, e) if e else []
        return acc.exp_var([base], [e])[0]

    def g_exp_prod(GG, arr, e):
        if GG is not G:
            return orig[1](GG, arr, e)
        return ar.gmap(lambda col: acc.expprod(col, e) if col else GG.one, arr)

    def g_mul(GG, a, b):
        if GG is not G:
            return orig[2](GG, a, b)
        return ar.gmap(lambda x, y: acc.mul(x, y) if isinstance(x, list) else GG.op_mul(x, y), a, b)

    ar.g_exp, ar.g_exp_prod, ar.g_mul = g_exp, g_exp_prod, g_mul

    def undo():
        ar.g_exp, ar.g_exp_prod, ar.g_mul = orig
    return undo
