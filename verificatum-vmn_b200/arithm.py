"""GPU-backed mirror of the `com.verificatum.arithm` array API the mix-net hot path calls.

Same class and method names as verificatum-vcr 3.1.0 (camelCase kept on purpose, so the
protocol code in hvzk.py / mixnet.py reads like the Java it mirrors and the Java call sites in
/root/reference map one to one):

    ModPGroup, PPGroup, PGroupElement, PPGroupElement, PGroupElementArray, PPGroupElementArray,
    PField (Z_q), PPRing, PFieldElement, PPRingElement, PRingElementArray, PPRingElementArray,
    LargeIntegerArray, Permutation

Arrays are opaque DEVICE handles of the C ABI (include/vmx.h); every array method is one C-ABI
call (or one per component of a product group).  Single elements are host values (Python ints),
exactly as single `PGroupElement`s are JVM-side `LargeInteger`s in the reference; their
exponentiations are routed through the engine as one-element arrays so that no group
arithmetic is ever done on the CPU.  Every temporary must be `free()`d, as in the reference.

Semantics that live in the un-vendored VCR jar ([VCR-mem] in SURVEY.md §8c) and are fixed here:
  * `permute(pi)`:  result[pi.map(i)] = this[i]
  * `Z_q.randomElementArray(n, rs, statDist)`: ceil((|q|+statDist)/8) bytes per element, masked
    to |q|+statDist bits, reduced mod q
  * `LargeIntegerArray.random(n, bits, rs)`: ceil(bits/8) bytes per element, masked to `bits`
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import List, Optional, Sequence

import numpy as np

from . import _native as nat
from .eio import (ByteTreeBasic, ByteTreeContainer, ByteTreeLeaf, ByteTreeLeafArray, ByteTreeReader, EIOException,
                  int_to_bytes)

ArithmFormatException = nat.ArithmFormatError


class ArithmError(RuntimeError):
    pass


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _be(x: int, width: int) -> bytes:
    return x.to_bytes(width, "big")


def _sha256_prg_offset(rs) -> Optional[int]:
    """Stream offset (bytes consumed so far) if `rs` is a PRGHeuristic(SHA-256): its expansion can
    then run on the device (counter mode) and the host object only accounts for the bytes; None
    for any other random source (its bytes are handed over)."""
    from .crypto import PRGHeuristic
    if isinstance(rs, PRGHeuristic) and rs.hf.name == "SHA-256" and rs.seed is not None and 32 <= len(rs.seed) <= 48:
        return rs.counter * 32 - len(rs.buf)
    return None


def _advance_prg(rs, offset: int) -> None:
    """Put the host-side PRG object at stream byte `offset`."""
    full, rem = divmod(offset, 32)
    rs.counter = full
    rs.buf = bytearray()
    if rem:
        rs.getBytes(rem)


def _jacobi(a: int, n: int) -> int:
    """Jacobi symbol (a|n), n odd positive (LargeInteger.legendre for prime n)."""
    a %= n
    t = 1
    while a:
        tz = (a & -a).bit_length() - 1
        a >>= tz
        if tz & 1 and n & 7 in (3, 5):
            t = -t
        if a & 3 == 3 and n & 3 == 3:
            t = -t
        a, n = n % a, a
    return t if n == 1 else 0


# ====================================================================== permutations
class Permutation:
    """com.verificatum.arithm.Permutation (table form)."""

    def __init__(self, table: Sequence[int]):
        self.table = np.ascontiguousarray(np.asarray(table, dtype=np.uint32))

    @staticmethod
    def identity(size: int) -> "Permutation":
        return Permutation(np.arange(size, dtype=np.uint32))

    @staticmethod
    def random(size: int, randomSource, statDist: int, group: Optional["ModPGroup"] = None) -> "Permutation":
        """Sort-based sampling: `size` integers of ceil(log2 size)+statDist bits, stable argsort.
        With `group` given and a PRGHeuristic(SHA-256) source, the random bytes are expanded on that
        group's device (vmx_prg_bytes_sha256) instead of one host hash per 32 bytes."""
        bits = max(1, (size - 1).bit_length()) + statDist
        nbytes = (bits + 7) // 8
        off = _sha256_prg_offset(randomSource) if group is not None else None
        if off is not None and size >= 4096:
            # keys drawn AND ranked on the device; only the table comes back
            table = np.empty(size, dtype=np.uint32)
            fb = C.c_int()
            nat.check(group._lib.vmx_permutation_prg_sha256(group.ctx, randomSource.seed, len(randomSource.seed), off,
                                                            size, nbytes, bits, _ptr(table), C.byref(fb)))
            if not fb.value:
                _advance_prg(randomSource, off + size * nbytes)
                return Permutation(table)
        if off is not None and size * nbytes >= 4096:
            raw = np.empty(size * nbytes, dtype=np.uint8)
            nat.check(group._lib.vmx_prg_bytes_sha256(group.ctx, randomSource.seed, len(randomSource.seed), off,
                                                      size * nbytes, _ptr(raw)))
            _advance_prg(randomSource, off + size * nbytes)
            raw = raw.reshape(size, nbytes)
        else:
            raw = np.frombuffer(randomSource.getBytes(size * nbytes), dtype=np.uint8).reshape(size, nbytes)
        # big-endian keys as columns of 64-bit words, most significant first; stable lexicographic sort
        pad = (-nbytes) % 8
        m = np.zeros((size, nbytes + pad), dtype=np.uint8)
        m[:, pad:] = raw
        m[:, pad] &= 0xFF >> ((8 - bits % 8) % 8)
        cols = m.view(">u8")
        # the keys are ~(log2 size + statDist)-bit random integers: sort by the leading 64-bit word and fall
        # back to the full lexicographic sort only if two leading words collide (same result either way)
        lead = cols[:, 0].astype(np.uint64)
        order = np.argsort(lead)  # distinct leading words: every sort gives the same order
        if size > 1:
            s_lead = lead[order]
            if (s_lead[1:] == s_lead[:-1]).any():
                order = np.lexsort([cols[:, j] for j in range(cols.shape[1] - 1, -1, -1)])
        table = np.empty(size, dtype=np.uint32)
        table[order] = np.arange(size, dtype=np.uint32)
        return Permutation(table)

    def size(self) -> int:
        return int(self.table.shape[0])

    def map(self, i: int) -> int:
        return int(self.table[i])

    def inv(self) -> "Permutation":
        """The inverse table (kept: the protocols ask for it several times per proof, hvzk/PoSBasicTW.java:552)."""
        if getattr(self, "_inv", None) is None:
            inv = np.empty_like(self.table)
            inv[self.table] = np.arange(self.table.shape[0], dtype=np.uint32)
            self._inv = Permutation(inv)
            self._inv._inv = self
        return self._inv

    def shrink(self, size: int) -> "Permutation":
        """Permutation.shrink (mixnet/PermutationCommitment.java:426): the restriction to the first
        `size` inputs, images renumbered 0..size-1 in increasing order ([VCR-mem]; it is the
        permutation under which commitment.extract(keepList) commits, keepList[map(i)] = true)."""
        img = self.table[:size].astype(np.int64)
        order = np.argsort(img, kind="stable")
        t = np.empty(size, dtype=np.uint32)
        t[order] = np.arange(size, dtype=np.uint32)
        return Permutation(t)

    def toByteTree(self) -> ByteTreeBasic:
        """One 4-byte leaf per entry ([VCR-mem]; only written to the prover's private state files)."""
        return ByteTreeContainer(*[ByteTreeLeaf(int(v).to_bytes(4, "big")) for v in self.table])

    def free(self) -> None:
        pass


_PINNED_MIN = 1 << 16


def _host_buffer(device: int, nbytes: int) -> np.ndarray:
    """Destination of a device-to-host copy: page-locked and pooled (vmx_host_alloc) when it is large enough
    to matter, so that the copy is one DMA transfer and does not hold up the stream behind it."""
    if nbytes < _PINNED_MIN:
        return np.empty(nbytes, dtype=np.uint8)
    try:
        return nat.host_buffer(device, nbytes)
    except nat.VmxError:      # the host cannot page-lock more memory: a pageable buffer is only slower
        return np.empty(nbytes, dtype=np.uint8)


def _install_host_buffers(device: int) -> None:
    """Published messages (ByteTreeBasic.to_buffer) are assembled in page-locked memory too: whoever imports
    them next (a verifier in the same process, a file writer) reads them by DMA."""
    from . import eio
    eio.set_buffer_factory(lambda nbytes: _host_buffer(device, nbytes))


class ByteTreeDeviceArray(ByteTreeBasic):
    """toByteTree() of a device array.  The serialisation -- the leaves of the node, headers included,
    written by the engine in exactly that form (vmx_*_to_leaves) -- is produced when the tree is first
    streamed (into a digest or a file) and cached ON THE ARRAY (arrays are immutable), so hashing an
    array that was just published, or that was imported from bytes, costs no second D2H copy."""

    def __init__(self, arr):
        self.arr = arr
        self._n = arr.size()
        self._w = arr.getPGroup().elem_bytes if hasattr(arr, "getPGroup") else arr.ring.byte_len
        # curve groups: node(node(x leaves), node(y leaves)); the engine writes the two inner nodes
        self._curve = hasattr(arr, "getPGroup") and getattr(arr.getPGroup(), "is_curve", False)

    def _stream(self) -> np.ndarray:
        if hasattr(self.arr, "leaves"):
            return self.arr.leaves()
        m = self.arr.to_matrix()  # sharded arrays: gathered matrix
        buf = np.empty((m.shape[0], 5 + self._w), dtype=np.uint8)
        buf[:, :5] = np.frombuffer(struct.pack(">BI", 1, self._w), dtype=np.uint8)
        buf[:, 5:] = m
        return buf.reshape(-1)

    def update(self, digest) -> None:
        digest.update(struct.pack(">BI", 0, 2 if self._curve else self._n))
        if not self._n and not self._curve:
            return
        arr = self.arr
        reserve = getattr(digest, "reserve", None)
        if reserve is not None and getattr(arr, "_leaves", 0) is None and getattr(arr, "h", None) is not None \
                and getattr(arr, "_export_into_messages", False):
            # a message is being assembled and this array has not been serialised yet: export into the message
            nbytes = arr.getPGroup()._leaves_bytes(self._n) if hasattr(arr, "getPGroup") else self._n * (5 + self._w)
            arr._export_leaves(reserve(nbytes))
            return
        digest.update(arr.leaves().data if self._curve else self._stream().data)

    def to_bytes(self) -> bytes:
        if self._curve:
            return struct.pack(">BI", 0, 2) + self.arr.leaves().tobytes()
        return struct.pack(">BI", 0, self._n) + (self._stream().tobytes() if self._n else b"")

    def total_bytes(self) -> int:
        if self._curve:
            return 5 + self.arr.getPGroup()._leaves_bytes(self._n)
        return 5 + self._n * (5 + self._w)


# ====================================================================== rings
class PRing:
    pass


class PField(PRing):
    """Z_q, the exponent ring of a prime-order group."""

    def __init__(self, group: "ModPGroup"):
        self.group = group
        self.order = group.q
        self.byte_len = group.ring_bytes

    def _rarr(self, h, size: Optional[int] = None) -> "PRingElementArray":
        """Wrap a fresh engine handle (overridden by the sharded field of parallel.py)."""
        return PRingElementArray(self, h)

    # -- elements
    def getZERO(self) -> "PFieldElement":
        return PFieldElement(self, 0)

    def getONE(self) -> "PFieldElement":
        return PFieldElement(self, 1)

    def getPField(self) -> "PField":
        return self

    def toElement(self, x) -> "PFieldElement":
        if isinstance(x, ByteTreeReader):
            if not x.isLeaf() or x.getRemaining() != self.byte_len:
                raise ArithmFormatException(nat.VMX_EFORMAT, "ring element of wrong length")
            v = int.from_bytes(x.read(), "big", signed=True)
            if not 0 <= v < self.order:
                raise ArithmFormatException(nat.VMX_EFORMAT, "ring element out of range")
            return PFieldElement(self, v)
        return PFieldElement(self, int(x) % self.order)

    def randomElement(self, randomSource, statDist: int) -> "PFieldElement":
        bits = self.order.bit_length() + statDist
        raw = randomSource.getBytes((bits + 7) // 8)
        return PFieldElement(self, (int.from_bytes(raw, "big") & ((1 << bits) - 1)) % self.order)

    # -- arrays
    def randomElementArray(self, size: int, randomSource, statDist: int) -> "PRingElementArray":
        bits = self.order.bit_length() + statDist
        width = (bits + 7) // 8
        off = _sha256_prg_offset(randomSource)
        if off is not None:
            h = C.c_void_p()
            nat.check(nat.load().vmx_rarr_prg_raw_sha256(self.group.ctx, randomSource.seed, len(randomSource.seed),
                                                         off, size, width, bits, C.byref(h)))
            _advance_prg(randomSource, off + size * width)
            return self._rarr(h, size)
        raw = np.frombuffer(randomSource.getBytes(size * width), dtype=np.uint8)
        return self._from_raw(size, raw, width, bits)

    def _from_raw(self, size, raw: np.ndarray, width: int, bits: int) -> "PRingElementArray":
        lib = nat.load()
        h = C.c_void_p()
        nat.check(lib.vmx_rarr_from_raw(self.group.ctx, size, _ptr(raw), width, bits, C.byref(h)))
        return self._rarr(h, size)

    def toElementArray(self, *args) -> "PRingElementArray":
        """(size, ByteTreeReader) | (LargeIntegerArray) | (size, element) | (list of elements)."""
        lib = nat.load()
        if len(args) == 1 and isinstance(args[0], LargeIntegerArray):
            return args[0]._to_ring(self)
        if len(args) == 1:
            vals = [int(e.value) for e in args[0]]
            m = np.frombuffer(b"".join(_be(v, self.byte_len) for v in vals), dtype=np.uint8)
            h = C.c_void_p()
            nat.check(lib.vmx_rarr_from_bytes(self.group.ctx, len(vals), _ptr(m), C.byref(h)))
            return self._rarr(h, len(vals))
        size, src = args
        h = C.c_void_p()
        if isinstance(src, ByteTreeReader):
            try:
                stream = src.leaf_stream(size, self.byte_len)
            except EIOException as e:
                raise ArithmFormatException(nat.VMX_EFORMAT, str(e))
            nat.check(lib.vmx_rarr_from_leaves(self.group.ctx, size, _ptr(stream), C.byref(h)))
            arr = self._rarr(h, size)
            arr._leaves = stream
            return arr
        else:
            nat.check(lib.vmx_rarr_fill(self.group.ctx, size, _be(src.value, self.byte_len), C.byref(h)))
        return self._rarr(h, size)

    unsafeToElementArray = toElementArray

    def __eq__(self, o):
        return isinstance(o, PField) and o.order == self.order

    def __hash__(self):
        return hash(self.order)


class PFieldElement:
    def __init__(self, ring: PField, value: int):
        self.ring = ring
        self.value = value % ring.order

    def getPRing(self):
        return self.ring

    def add(self, o):
        return PFieldElement(self.ring, self.value + o.value)

    def sub(self, o):
        return PFieldElement(self.ring, self.value - o.value)

    def mul(self, o):
        return PFieldElement(self.ring, self.value * o.value)

    def neg(self):
        return PFieldElement(self.ring, -self.value)

    def mulAdd(self, v, b):
        """this * v + b (hvzk/PoSBasicTW.java:873-878)."""
        return PFieldElement(self.ring, self.value * v.value + b.value)

    def toLargeInteger(self) -> int:
        return self.value

    def toByteTree(self) -> ByteTreeBasic:
        return ByteTreeLeaf(int_to_bytes(self.value, self.ring.byte_len))

    def equals(self, o) -> bool:
        return isinstance(o, PFieldElement) and o.ring == self.ring and o.value == self.value

    __eq__ = equals

    def __hash__(self):
        return hash(self.value)


class PRingElementArray:
    """Device-resident array over Z_q."""

    _export_into_messages = True   # ByteTreeDeviceArray.update: serialise straight into a message being assembled

    def __init__(self, ring: PField, handle):
        self.ring = ring
        self.h = handle
        self._lib = nat.load()
        self._leaves = None  # cached serialisation (leaves of the byte tree)

    # -- bookkeeping
    def getPRing(self):
        return self.ring

    def size(self) -> int:
        return int(self._lib.vmx_rarr_size(self.h))

    def free(self) -> None:
        if self.h is not None and self.h.value:
            self._lib.vmx_rarr_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _new(self, h) -> "PRingElementArray":
        return PRingElementArray(self.ring, h)

    def _scalar_out(self, fn, *args) -> PFieldElement:
        buf = np.empty(self.ring.byte_len, dtype=np.uint8)
        nat.check(fn(*args, _ptr(buf)))
        return PFieldElement(self.ring, int.from_bytes(buf.tobytes(), "big"))

    # -- algebra
    def add(self, o):
        h = C.c_void_p()
        nat.check(self._lib.vmx_radd(self.h, o.h, C.byref(h)))
        return self._new(h)

    def neg(self):
        h = C.c_void_p()
        nat.check(self._lib.vmx_rneg(self.h, C.byref(h)))
        return self._new(h)

    def mul(self, o):
        h = C.c_void_p()
        if isinstance(o, PFieldElement):
            zero = self.ring.toElementArray(self.size(), self.ring.getZERO())
            try:
                return self.mulAdd(o, zero)
            finally:
                zero.free()
        nat.check(self._lib.vmx_rmul(self.h, o.h, C.byref(h)))
        return self._new(h)

    def mulAdd(self, scalar: PFieldElement, o: "PRingElementArray"):
        """this[i] * scalar + o[i] (hvzk/PoSBasicTW.java:874,877)."""
        h = C.c_void_p()
        nat.check(self._lib.vmx_rmuladd(self.h, _be(scalar.value, self.ring.byte_len), o.h, C.byref(h)))
        return self._new(h)

    def innerProduct(self, o) -> PFieldElement:
        return self._scalar_out(self._lib.vmx_rinner, self.h, o.h)

    def sum(self) -> PFieldElement:
        return self._scalar_out(self._lib.vmx_rsum, self.h)

    def prod(self) -> PFieldElement:
        return self._scalar_out(self._lib.vmx_rprod, self.h)

    def prods(self):
        h = C.c_void_p()
        nat.check(self._lib.vmx_rprods(self.h, C.byref(h)))
        return self._new(h)

    def recLin(self, e: "PRingElementArray"):
        """x[0] = this[0], x[i] = x[i-1]*e[i] + this[i]; returns (x, x[n-1]) (PoSBasicTW.java:583-598)."""
        h = C.c_void_p()
        buf = np.empty(self.ring.byte_len, dtype=np.uint8)
        nat.check(self._lib.vmx_rreclin(self.h, e.h, C.byref(h), _ptr(buf)))
        return self._new(h), PFieldElement(self.ring, int.from_bytes(buf.tobytes(), "big"))

    # -- data movement
    def permute(self, pi: Permutation):
        h = C.c_void_p()
        nat.check(self._lib.vmx_rpermute(self.h, _ptr(pi.table), C.byref(h)))
        return self._new(h)

    def shiftPush(self, el: PFieldElement):
        h = C.c_void_p()
        nat.check(self._lib.vmx_rshift_push(self.h, _be(el.value, self.ring.byte_len), C.byref(h)))
        return self._new(h)

    def copyOfRange(self, a: int, b: int):
        h = C.c_void_p()
        nat.check(self._lib.vmx_rslice(self.h, a, b, C.byref(h)))
        return self._new(h)

    def get(self, i: int) -> PFieldElement:
        return self._scalar_out(lambda hh, ii, out: self._lib.vmx_rget(hh, ii, out), self.h, i)

    def equals(self, o) -> bool:
        eq = C.c_int()
        nat.check(self._lib.vmx_requals(self.h, o.h, C.byref(eq)))
        return bool(eq.value)

    def bitLength(self) -> int:
        b = C.c_uint()
        nat.check(self._lib.vmx_rarr_bitlen(self.h, C.byref(b)))
        return int(b.value)

    # -- I/O
    def leaves(self) -> np.ndarray:
        """The n * (5 + width) bytes this array serialises to (leaf headers included), cached."""
        if self._leaves is None:
            if self.h is None:
                raise ArithmError("byte tree of a freed array")
            buf = _host_buffer(self.ring.group.device, self.size() * (5 + self.ring.byte_len))
            self._export_leaves(buf)
        return self._leaves

    def _export_leaves(self, buf: np.ndarray) -> None:
        """Serialise into `buf` (n * (5 + width) bytes) and keep it as this array's cached serialisation.  A message
        being assembled (eio._Writer.reserve) hands its own slice in: the engine then exports straight into the
        published message -- no second copy of the array's 390 MB per 10^6 elements."""
        nat.check(self._lib.vmx_rarr_to_leaves(self.h, _ptr(buf)))
        self._leaves = buf

    def to_matrix(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        n, w = self.size(), self.ring.byte_len
        m = self.leaves().reshape(n, 5 + w)[:, 5:]
        if out is not None:
            out[:] = m
            return out
        return m

    def toByteTree(self) -> ByteTreeBasic:
        return ByteTreeDeviceArray(self)

    def elements(self) -> List[PFieldElement]:
        return [PFieldElement(self.ring, int.from_bytes(r.tobytes(), "big")) for r in self.to_matrix()]

    @staticmethod
    def free_(a) -> None:
        if a is not None:
            a.free()


class LargeIntegerArray:
    """Non-negative integers of bounded bit length, device-resident (already < q here)."""

    def __init__(self, field: PField, handle, size: Optional[int] = None):
        self.field = field
        self.h = handle
        self.gsize = size

    @staticmethod
    def random(size: int, bitLength: int, randomSource, field: PField) -> "LargeIntegerArray":
        """LargeIntegerArray.random(size, bitLength, randomSource) (PoSBasicTW.java:472-474,535-536).

        With a PRGHeuristic(SHA-256) source the expansion runs on the device (counter mode);
        any other source hands its bytes over."""
        if hasattr(field, "_lia_random"):  # sharded field (parallel.py): every rank draws its own slice
            return field._lia_random(size, bitLength, randomSource)
        lib = nat.load()
        h = C.c_void_p()
        width = (bitLength + 7) // 8
        off = _sha256_prg_offset(randomSource)
        if off is not None and bitLength < field.order.bit_length():
            nat.check(lib.vmx_rarr_prg_sha256(field.group.ctx, randomSource.seed, len(randomSource.seed), off, size,
                                              bitLength, C.byref(h)))
            _advance_prg(randomSource, off + size * width)
            return LargeIntegerArray(field, h)
        if off is not None:
            # integers as wide as or wider than q (n_e + n_v + n_r = 612 bits over the 256-bit order of a curve
            # group): every use converts them to field elements (pField.toElementArray, hvzk/PoSBasicTW.java:473),
            # so they are reduced mod q as they are drawn
            nat.check(lib.vmx_rarr_prg_raw_sha256(field.group.ctx, randomSource.seed, len(randomSource.seed), off, size,
                                                  width, bitLength, C.byref(h)))
            _advance_prg(randomSource, off + size * width)
            return LargeIntegerArray(field, h)
        raw = np.frombuffer(randomSource.getBytes(size * width), dtype=np.uint8)
        nat.check(lib.vmx_rarr_from_raw(field.group.ctx, size, _ptr(raw), width, bitLength, C.byref(h)))
        return LargeIntegerArray(field, h)

    def _to_ring(self, field: PField) -> PRingElementArray:
        h, self.h = self.h, None
        return field._rarr(h, self.gsize)

    def free(self) -> None:
        if self.h is not None and self.h.value:
            nat.load().vmx_rarr_free(self.h)
        self.h = None


# ====================================================================== groups
class PGroup:
    pass


class ModPGroup(PGroup):
    """Subgroup of order q of Z_p^* with generator g (com.verificatum.arithm.ModPGroup)."""

    def __init__(self, p: int, q: int, g: int, device: int = 0):
        lib = nat.load()
        self.p, self.q, self.g = p, q, g
        nbytes = (p.bit_length() + 7) // 8
        ctx = C.c_void_p()
        nat.check(lib.vmx_ctx_create_modp(_be(p, nbytes), _be(q, nbytes), _be(g, nbytes), nbytes, device,
                                          C.byref(ctx)))
        self.ctx = ctx
        self._lib = lib
        self.device = device
        _install_host_buffers(device)
        self.elem_bytes = int(lib.vmx_ctx_elem_bytes(ctx))
        self.ring_bytes = int(lib.vmx_ctx_ring_bytes(ctx))
        self.pRing = PField(self)

    def __del__(self):
        # arrays keep a reference to their group, so the context outlives every handle
        try:
            import sys
            if sys.is_finalizing():
                return  # interpreter teardown order is arbitrary; the process exit frees the device
            if self.ctx is not None:
                self._lib.vmx_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def _garr(self, h, size: Optional[int] = None) -> "PGroupElementArray":
        """Wrap a fresh engine handle (overridden by the sharded group of parallel.py)."""
        return PGroupElementArray(self, h)

    is_curve = False

    def _leaves_bytes(self, n: int) -> int:
        """Size of the buffer vmx_garr_to_leaves / vmx_garr_from_leaves move for n elements."""
        return n * (5 + self.elem_bytes)

    def _elem_tree(self, value: int) -> ByteTreeBasic:
        return ByteTreeLeaf(int_to_bytes(value, self.elem_bytes))

    def _combine_partials(self, parts: List["PGroupElement"]) -> List["PGroupElement"]:
        """expProd / prod results of this process; the sharded group multiplies the ranks' partial
        products here (parallel.py)."""
        return parts

    # -- structure
    def getPRing(self) -> PField:
        return self.pRing

    def getg(self) -> "PGroupElement":
        return PGroupElement(self, self.g)

    def getONE(self) -> "PGroupElement":
        return PGroupElement(self, 1)

    def getElementOrder(self) -> int:
        return self.q

    def basic(self) -> List["ModPGroup"]:
        return [self]

    def launch_count(self) -> int:
        return int(self._lib.vmx_ctx_launch_count(self.ctx))

    def modmul_count(self) -> int:
        return int(self._lib.vmx_ctx_modmul_count(self.ctx))

    def sync(self) -> None:
        nat.check(self._lib.vmx_ctx_sync(self.ctx))

    def set_tuning(self, **knobs: int) -> None:
        """Kernel-selection knobs of the engine (vmx_ctx_set_tuning: coop_max, var_chunk, mexp_window,
        fixed_window); none changes a result."""
        for k, v in knobs.items():
            nat.check(self._lib.vmx_ctx_set_tuning(self.ctx, k.encode(), int(v)))

    def precomputeFixedBase(self, el: "PGroupElement", size_hint: int) -> None:
        nat.check(self._lib.vmx_fixed_precompute(self.ctx, _be(el.value, self.elem_bytes), size_hint))

    # -- elements
    def toElement(self, x) -> "PGroupElement":
        if isinstance(x, ByteTreeReader):
            if not x.isLeaf() or x.getRemaining() != self.elem_bytes:
                raise ArithmFormatException(nat.VMX_EFORMAT, "group element of wrong length")
            v = int.from_bytes(x.read(), "big", signed=True)
        else:
            v = int(x)
        if not 0 < v < self.p:
            raise ArithmFormatException(nat.VMX_EFORMAT, "group element out of range")
        if self.p == 2 * self.q + 1:
            # safe prime: the order-q subgroup is the set of quadratic residues; a single
            # element is validated by its Legendre symbol, as VCR does for one LargeInteger
            if _jacobi(v, self.p) != 1:
                raise ArithmFormatException(nat.VMX_EFORMAT, "element not in the order-q subgroup")
        else:
            # general subgroup: Euler criterion on the device (one-element array)
            m = np.frombuffer(_be(v, self.elem_bytes), dtype=np.uint8)
            h = C.c_void_p()
            nat.check(self._lib.vmx_garr_from_bytes(self.ctx, 1, _ptr(m), 1, C.byref(h)))
            self._lib.vmx_garr_free(h)
        return PGroupElement(self, v)

    # -- arrays
    # import-time subgroup check of untrusted arrays (PGroup.toElementArray in VCR always checks)
    membership_check = True

    def toElementArray(self, *args, check_membership: Optional[bool] = None) -> "PGroupElementArray":
        """(size, ByteTreeReader) | (size, PGroupElement) | (list of PGroupElement)."""
        lib = self._lib
        if check_membership is None:
            check_membership = self.membership_check
        h = C.c_void_p()
        if len(args) == 1:
            vals = [e.value for e in args[0]]
            m = np.frombuffer(b"".join(_be(v, self.elem_bytes) for v in vals), dtype=np.uint8)
            nat.check(lib.vmx_garr_from_bytes(self.ctx, len(vals), _ptr(m), 0, C.byref(h)))
            return self._garr(h, len(vals))
        size, src = args
        if isinstance(src, ByteTreeReader):
            try:
                stream = src.leaf_stream(size, self.elem_bytes)
            except EIOException as e:
                raise ArithmFormatException(nat.VMX_EFORMAT, str(e))
            nat.check(lib.vmx_garr_from_leaves(self.ctx, size, _ptr(stream), 1 if check_membership else 0, C.byref(h)))
            arr = self._garr(h, size)
            arr._leaves = stream
            return arr
        elif isinstance(src, np.ndarray):
            src = np.ascontiguousarray(src)
            nat.check(lib.vmx_garr_from_bytes(self.ctx, size, _ptr(src), 1 if check_membership else 0, C.byref(h)))
        else:
            nat.check(lib.vmx_garr_fill(self.ctx, size, _be(src.value, self.elem_bytes), C.byref(h)))
        return self._garr(h, size)

    def unsafeToElementArray(self, *args) -> "PGroupElementArray":
        return self.toElementArray(*args, check_membership=False)

    def randomElementArray(self, size: int, randomSource, statDist: int) -> "PGroupElementArray":
        """t_i = next ceil((|p|+statDist)/8) bytes mod 2^(|p|+statDist); h_i = t_i^((p-1)/q) mod p
        (distr/IndependentGeneratorsRO.java:129).  The reduction mod p and the cofactor power run
        on the device."""
        bits = self.p.bit_length() + statDist
        width = (bits + 7) // 8
        h = C.c_void_p()
        off = _sha256_prg_offset(randomSource)
        if off is not None:
            nat.check(self._lib.vmx_garr_prg_sha256(self.ctx, randomSource.seed, len(randomSource.seed), off, size,
                                                    width, bits, C.byref(h)))
            _advance_prg(randomSource, off + size * width)
            return self._garr(h, size)
        raw = np.frombuffer(randomSource.getBytes(size * width), dtype=np.uint8)
        nat.check(self._lib.vmx_garr_from_raw(self.ctx, size, _ptr(raw), width, bits, C.byref(h)))
        return self._garr(h, size)

    def expProd(self, bases: Sequence["PGroupElementArray"], integers: Sequence[int], bitLength: int):
        """Element-wise prod_j bases[j][i]^integers[j] (elgamal/DistrElGamalSessionBasic.java:502)."""
        t = len(bases)
        arr = (C.c_void_p * t)(*[b.h for b in bases])
        ints = (C.c_int64 * t)(*[int(x) for x in integers])
        h = C.c_void_p()
        nat.check(self._lib.vmx_expprod_cols(arr, t, ints, C.byref(h)))
        return self._garr(h, bases[0].size())

    def __eq__(self, o):
        return isinstance(o, ModPGroup) and (o.p, o.q, o.g) == (self.p, self.q, self.g)

    def __hash__(self):
        return hash((self.p, self.g))



class ECqPGroup(ModPGroup):
    """Prime-order elliptic-curve group y^2 = x^3 + ax + b over F_p (com.verificatum.arithm.ECqPGroup;
    the reference's default groups are NIST curves: demo/mixnet/benchmarks/bench_config:33-50, P-256).

    A single element is held as the integer of its wire form x || y (two fixed-width two's-complement
    coordinates, the unit element (-1, -1)), so that PGroupElement and the array classes above serve both
    kinds of group; only the byte-tree framing differs ([VCR-mem]):
    element = node(leaf(x), leaf(y)), array = node(node(x_0..), node(y_0..))."""

    is_curve = True
    CURVES = {
        "P-256": dict(
            p=0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF,
            a=0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFC,
            b=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
            gx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
            gy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
            n=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551),
        "secp256k1": dict(
            p=2 ** 256 - 2 ** 32 - 977, a=0, b=7,
            gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
            gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
            n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141),
    }

    def __init__(self, name: str = "P-256", device: int = 0, **params):
        lib = nat.load()
        c = dict(self.CURVES[name]) if name in self.CURVES else {}
        c.update(params)
        self.name = name
        self.p, self.a, self.b, self.q = c["p"], c["a"] % c["p"], c["b"] % c["p"], c["n"]
        self.gx, self.gy = c["gx"], c["gy"]
        ctx = C.c_void_p()
        nat.check(lib.vmx_ctx_create_ecq(_be(self.p, 32), _be(self.a, 32), _be(self.b, 32), _be(self.gx, 32),
                                         _be(self.gy, 32), _be(self.q, 32), 32, device, C.byref(ctx)))
        self.ctx = ctx
        self._lib = lib
        self.device = device
        _install_host_buffers(device)
        self.elem_bytes = int(lib.vmx_ctx_elem_bytes(ctx))
        self.coord_bytes = self.elem_bytes // 2
        self.ring_bytes = int(lib.vmx_ctx_ring_bytes(ctx))
        self.pRing = PField(self)
        self.g = self._pack(self.gx, self.gy)
        self.one = (1 << (8 * self.elem_bytes)) - 1  # (-1, -1)

    # -- wire form of one point
    def _pack(self, x: int, y: int) -> int:
        return int.from_bytes(int_to_bytes(x, self.coord_bytes) + int_to_bytes(y, self.coord_bytes), "big")

    def _unpack(self, value: int):
        raw = _be(value, self.elem_bytes)
        cb = self.coord_bytes
        return int.from_bytes(raw[:cb], "big", signed=True), int.from_bytes(raw[cb:], "big", signed=True)

    def _leaves_bytes(self, n: int) -> int:
        return 2 * (5 + n * (5 + self.coord_bytes))

    def _elem_tree(self, value: int) -> ByteTreeBasic:
        raw = _be(value, self.elem_bytes)
        cb = self.coord_bytes
        return ByteTreeContainer(ByteTreeLeaf(raw[:cb]), ByteTreeLeaf(raw[cb:]))

    def getONE(self) -> "PGroupElement":
        return PGroupElement(self, self.one)

    def _on_curve(self, x: int, y: int) -> bool:
        return 0 <= x < self.p and 0 <= y < self.p and (y * y - (x * x * x + self.a * x + self.b)) % self.p == 0

    def toElement(self, x) -> "PGroupElement":
        cb = self.coord_bytes
        if isinstance(x, ByteTreeReader):
            if x.isLeaf() or x.getRemaining() != 2:
                raise ArithmFormatException(nat.VMX_EFORMAT, "curve point of wrong arity")
            cx, cy = x.getNextChild(), x.getNextChild()
            if not cx.isLeaf() or not cy.isLeaf() or cx.getRemaining() != cb or cy.getRemaining() != cb:
                raise ArithmFormatException(nat.VMX_EFORMAT, "coordinate of wrong length")
            px, py = int.from_bytes(cx.read(), "big", signed=True), int.from_bytes(cy.read(), "big", signed=True)
        else:
            px, py = x
        if (px, py) == (-1, -1):
            return self.getONE()
        # a single point is validated on the host, as VCR does for one ECqPGroupElement (a handful of
        # field operations on host integers; arrays are checked on the device)
        if not self._on_curve(px, py):
            raise ArithmFormatException(nat.VMX_EFORMAT, "not a point of the curve")
        return PGroupElement(self, self._pack(px, py))

    def toElementArray(self, *args, check_membership: Optional[bool] = None) -> "PGroupElementArray":
        lib = self._lib
        h = C.c_void_p()
        if len(args) == 2 and isinstance(args[1], ByteTreeReader):
            size, src = args
            try:
                stream = src.point_array_stream(size, self.coord_bytes)
            except EIOException as e:
                raise ArithmFormatException(nat.VMX_EFORMAT, str(e))
            nat.check(lib.vmx_garr_from_leaves(self.ctx, size, _ptr(stream), 1, C.byref(h)))
            arr = self._garr(h, size)
            arr._leaves = stream
            return arr
        return super().toElementArray(*args, check_membership=check_membership)

    def _random_full_handle(self, size: int, randomSource, statDist: int):
        """Engine handle of `size` random points drawn from `randomSource` (the whole array on this device)."""
        bits = self.p.bit_length() + statDist
        width = (bits + 7) // 8
        h = C.c_void_p()
        off = _sha256_prg_offset(randomSource)
        if off is not None:
            nat.check(self._lib.vmx_garr_prg_sha256(self.ctx, randomSource.seed, len(randomSource.seed), off, size,
                                                    width, bits, C.byref(h)))
            _advance_prg(randomSource, off + int(self._lib.vmx_ctx_prg_consumed(self.ctx)))
            return h
        # any other source: its bytes are consumed candidate by candidate, so ask for exactly one candidate
        # per missing point and round (about log2(size) rounds)
        rows, have = [], 0
        while have < size:
            m = size - have
            raw = np.frombuffer(randomSource.getBytes(m * width), dtype=np.uint8)
            used = C.c_size_t()
            hh = C.c_void_p()
            nat.check(self._lib.vmx_garr_from_candidates(self.ctx, m, _ptr(raw), width, bits, m, C.byref(hh),
                                                         C.byref(used)))
            part = PGroupElementArray(self, hh)
            have += part.size()
            rows.append(part.to_matrix())
            part.free()
        m = np.ascontiguousarray(np.concatenate(rows)) if rows else np.empty((0, self.elem_bytes), dtype=np.uint8)
        nat.check(self._lib.vmx_garr_from_bytes(self.ctx, size, _ptr(m), 0, C.byref(h)))
        return h

    def randomElementArray(self, size: int, randomSource, statDist: int) -> "PGroupElementArray":
        """Per point: draw ceil((|p|+statDist)/8) bytes, reduce mod p to x, accept if x^3+ax+b is a square and
        take the smaller root, else draw again ([VCR-mem]; distr/IndependentGeneratorsRO.java:129).  The
        candidates are tested in parallel on the device and compacted in stream order."""
        return self._garr(self._random_full_handle(size, randomSource, statDist), size)

    def __eq__(self, o):
        return isinstance(o, ECqPGroup) and (o.p, o.a, o.b, o.g, o.q) == (self.p, self.a, self.b, self.g, self.q)

    def __hash__(self):
        return hash((self.p, self.b, self.g))


class PGroupElement:
    """A single group element: host value; exponentiations go through the engine."""

    def __init__(self, group: ModPGroup, value: int):
        self.group = group
        self.value = value

    def getPGroup(self):
        return self.group

    def _be(self) -> bytes:
        return _be(self.value, self.group.elem_bytes)

    def _single(self, h) -> "PGroupElement":
        lib = self.group._lib
        buf = np.empty(self.group.elem_bytes, dtype=np.uint8)
        try:
            nat.check(lib.vmx_get(h, 0, _ptr(buf)))
        finally:
            lib.vmx_garr_free(h)
        return PGroupElement(self.group, int.from_bytes(buf.tobytes(), "big"))

    def _as_array(self):
        h = C.c_void_p()
        nat.check(self.group._lib.vmx_garr_fill(self.group.ctx, 1, self._be(), C.byref(h)))
        return h

    def exp(self, e):
        """exp(PRingElementArray) -> fixed-base array exponentiation (ShufflerElGamalSession.java:407;
        PoSBasicTW.java:447,606,608,644,646,1030); exp(PRingElement) -> single element."""
        lib = self.group._lib
        if isinstance(e, PRingElementArray):
            h = C.c_void_p()
            nat.check(lib.vmx_exp_fixed(self.group.ctx, self._be(), e.h, C.byref(h)))
            return self.group._garr(h, e.size())
        if isinstance(e, int):
            e = self.group.pRing.toElement(e)
        # An exponent just below the group order is a small negative one (the modified Lagrange coefficients of
        # DistrElGamalSessionBasic.combine, :642-678, arrive as q - |lambda|): x^(q-m) = (x^-1)^m for x of order
        # dividing q -- an inversion and a short chain instead of |q| squarings on one warp.
        x, e = self._short_form(e)
        if x is not self:
            return x.exp(e)
        buf = np.empty(self.group.elem_bytes, dtype=np.uint8)
        nat.check(lib.vmx_elem_exp(self.group.ctx, self._be(), _be(e.value, self.group.ring_bytes), _ptr(buf)))
        return PGroupElement(self.group, int.from_bytes(buf.tobytes(), "big"))

    def _short_form(self, e: PFieldElement):
        """(x, e) -> (x^-1, q - e) when that exponent is much shorter and x^q = 1; else unchanged."""
        m = self.group.q - e.value
        if e.value and m.bit_length() + 64 < e.value.bit_length() and self._order_divides_q():
            return self.inv(), self.group.pRing.toElement(m)
        return self, e

    def _order_divides_q(self) -> bool:
        G = self.group
        if G.is_curve:
            return True     # prime-order curve: every point of the group
        if G.p == 2 * G.q + 1:
            return _jacobi(self.value, G.p) == 1
        return False

    def mul(self, o: "PGroupElement") -> "PGroupElement":
        lib = self.group._lib
        a, b = self._as_array(), o._as_array()
        try:
            h = C.c_void_p()
            nat.check(lib.vmx_mul(a, b, C.byref(h)))
        finally:
            lib.vmx_garr_free(a)
            lib.vmx_garr_free(b)
        return self._single(h)

    def inv(self) -> "PGroupElement":
        buf = np.empty(self.group.elem_bytes, dtype=np.uint8)
        nat.check(self.group._lib.vmx_elem_inv(self.group.ctx, self._be(), _ptr(buf)))
        return PGroupElement(self.group, int.from_bytes(buf.tobytes(), "big"))

    def div(self, o: "PGroupElement") -> "PGroupElement":
        return self.mul(o.inv())

    def expMul(self, e: PFieldElement, b: "PGroupElement") -> "PGroupElement":
        """this^e * b (PoSBasicTW.java:1021,1048,1055,1063)."""
        return self.exp(e).mul(b)

    def equals(self, o) -> bool:
        return isinstance(o, PGroupElement) and o.group == self.group and o.value == self.value

    __eq__ = equals

    def __hash__(self):
        return hash(self.value)

    def toByteTree(self) -> ByteTreeBasic:
        return self.group._elem_tree(self.value)


def expMany(elements: Sequence["PGroupElement"], exponents: Sequence[PFieldElement]) -> List["PGroupElement"]:
    """[x_j^{e_j}] for a handful of unrelated single elements (the Sigma-protocol of the decryption proof,
    elgamal/DistrElGamalSessionBasic.java:642-727).  One exponentiation of a single element is a chain of |e|
    dependent squarings on one warp (about 21 ms at 3072 bits); as ONE small array they run side by side, one
    warp each.  Elements of product groups and of sharded groups (replicated on every rank) take the loop."""
    elements, exponents = list(elements), list(exponents)
    G = elements[0].group if elements else None
    plain = all(type(x) is PGroupElement and x.group is G for x in elements) and type(G) in (ModPGroup, ECqPGroup)
    if len(elements) < 2 or not plain:
        return [x.exp(e) for x, e in zip(elements, exponents)]
    pairs = [x._short_form(e) for x, e in zip(elements, exponents)]
    X = G.toElementArray([x for x, _ in pairs])
    E = G.pRing.toElementArray([e for _, e in pairs])
    R = X.exp(E)
    out = R.elements()
    X.free()
    E.free()
    R.free()
    return out


class PGroupElementArray:
    """Device-resident array of ModPGroup elements (Montgomery form, limb-major)."""

    _export_into_messages = True   # ByteTreeDeviceArray.update: serialise straight into a message being assembled

    def __init__(self, group: ModPGroup, handle):
        self.group = group
        self.h = handle
        self._lib = group._lib
        self._leaves = None  # cached serialisation (leaves of the byte tree)

    def getPGroup(self):
        return self.group

    def size(self) -> int:
        return int(self._lib.vmx_garr_size(self.h))

    def free(self) -> None:
        if self.h is not None and self.h.value:
            self._lib.vmx_garr_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _new(self, h):
        return PGroupElementArray(self.group, h)

    def basic(self) -> List["PGroupElementArray"]:
        return [self]

    # -- algebra
    def mul(self, o: "PGroupElementArray"):
        h = C.c_void_p()
        nat.check(self._lib.vmx_mul(self.h, o.h, C.byref(h)))
        return self._new(h)

    def inv(self):
        h = C.c_void_p()
        nat.check(self._lib.vmx_inv(self.h, C.byref(h)))
        return self._new(h)

    def exp(self, e):
        """exp(PRingElementArray): per-element exponents (PoSBasicTW.java:1032);
        exp(PRingElement): one exponent for all (PoSBasicTW.java:1028; DistrElGamalSession.java:384)."""
        h = C.c_void_p()
        if isinstance(e, PRingElementArray):
            nat.check(self._lib.vmx_exp_var(self.h, e.h, C.byref(h)))
        else:
            nat.check(self._lib.vmx_exp_scalar(self.h, _be(e.value, self.group.ring_bytes), C.byref(h)))
        return self._new(h)

    def expMulExp(self, x: PFieldElement, other: "PGroupElementArray", y: PRingElementArray):
        """this[i]^x * other[i]^y[i] by simultaneous exponentiation (vmx_exp_scalar_var): what the verifier's
        B.exp(v) and B_shift.exp(k_E) (PoSBasicTW.java:1028-1032) become once both are on one side."""
        h = C.c_void_p()
        nat.check(self._lib.vmx_exp_scalar_var(self.h, _be(x.value, self.group.ring_bytes), other.h, y.h, C.byref(h)))
        return self._new(h)

    def expProd(self, e: PRingElementArray) -> PGroupElement:
        """prod_i this[i]^e[i] (PoSBasicTW.java:408-409,481,690,1021,1063)."""
        return expProdMany([self], e)[0]

    def prod(self) -> PGroupElement:
        buf = np.empty(self.group.elem_bytes, dtype=np.uint8)
        nat.check(self._lib.vmx_prod(self.h, _ptr(buf)))
        return self.group._combine_partials([PGroupElement(self.group, int.from_bytes(buf.tobytes(), "big"))])[0]

    # -- data movement
    def permute(self, pi: Permutation):
        h = C.c_void_p()
        nat.check(self._lib.vmx_permute(self.h, _ptr(pi.table), C.byref(h)))
        return self._new(h)

    def shiftPush(self, el: PGroupElement):
        h = C.c_void_p()
        nat.check(self._lib.vmx_shift_push(self.h, el._be(), C.byref(h)))
        return self._new(h)

    def extract(self, keep: Sequence[bool]):
        k = np.ascontiguousarray(np.asarray(keep, dtype=np.uint8))
        h = C.c_void_p()
        nat.check(self._lib.vmx_extract(self.h, k.ctypes.data_as(C.c_char_p), C.byref(h)))
        return self._new(h)

    def copyOfRange(self, a: int, b: int):
        h = C.c_void_p()
        nat.check(self._lib.vmx_slice(self.h, a, b, C.byref(h)))
        return self._new(h)

    def get(self, i: int) -> PGroupElement:
        buf = np.empty(self.group.elem_bytes, dtype=np.uint8)
        nat.check(self._lib.vmx_get(self.h, i, _ptr(buf)))
        return PGroupElement(self.group, int.from_bytes(buf.tobytes(), "big"))

    def equals(self, o) -> bool:
        eq = C.c_int()
        nat.check(self._lib.vmx_equals(self.h, o.h, C.byref(eq)))
        return bool(eq.value)

    # -- I/O
    def leaves(self) -> np.ndarray:
        """The n * (5 + width) bytes this array serialises to (leaf headers included), cached."""
        if self._leaves is None:
            if self.h is None:
                raise ArithmError("byte tree of a freed array")
            buf = _host_buffer(self.group.device, self.group._leaves_bytes(self.size()))
            self._export_leaves(buf)
        return self._leaves

    def _export_leaves(self, buf: np.ndarray) -> None:
        """As PRingElementArray._export_leaves."""
        nat.check(self._lib.vmx_garr_to_leaves(self.h, _ptr(buf)))
        self._leaves = buf

    def to_matrix(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        n, w = self.size(), self.group.elem_bytes
        if self.group.is_curve:  # x || y per point
            m = np.empty((n, w), dtype=np.uint8)
            if n:
                nat.check(self._lib.vmx_garr_to_bytes(self.h, _ptr(m)))
            if out is not None:
                out[:] = m
                return out
            return m
        m = self.leaves().reshape(n, 5 + w)[:, 5:]
        if out is not None:
            out[:] = m
            return out
        return m

    def toByteTree(self) -> ByteTreeBasic:
        return ByteTreeDeviceArray(self)

    def elements(self) -> List[PGroupElement]:
        return [PGroupElement(self.group, int.from_bytes(r.tobytes(), "big")) for r in self.to_matrix()]

    @staticmethod
    def free_(a) -> None:
        if a is not None:
            a.free()


def expProdMany(arrays: Sequence[PGroupElementArray], e: PRingElementArray) -> List[PGroupElement]:
    """expProd of several arrays that share one exponent array: the engine sorts the exponent
    digits once (the components of a product-group array, PoSBasicTW.java:409,690,1063)."""
    group = arrays[0].group
    k = len(arrays)
    arr = (C.c_void_p * k)(*[a.h for a in arrays])
    buf = np.empty((k, group.elem_bytes), dtype=np.uint8)
    nat.check(group._lib.vmx_expprod(arr, k, e.h, _ptr(buf)))
    return group._combine_partials([PGroupElement(group, int.from_bytes(buf[i].tobytes(), "big")) for i in range(k)])


def expProdTogether(arrays: Sequence, e: PRingElementArray) -> list:
    """[a.expProd(e) for a in arrays] for arrays (of a basic or of a product group) that share the exponent
    array e -- u and w in computeAF (PoSBasicTW.java:408-409), h and w' in the checks (:1021,1063): ONE call into
    the engine, so the exponent digits are sorted once and the Horner chains of all results (|e| dependent
    squarings each, on one warp) run side by side instead of one after the other."""
    flats = [a.basic() for a in arrays]
    it = iter(expProdMany([x for f in flats for x in f], e))

    def rebuild(arr):
        if isinstance(arr, PPGroupElementArray):
            return PPGroupElement(arr.group, [rebuild(c) for c in arr.comps])
        return next(it)
    return [rebuild(a) for a in arrays]


# ====================================================================== product groups / rings
class PPRing(PRing):
    """Product ring: the exponent ring of a product group (one factor ring per group factor)."""

    def __init__(self, factors, width: Optional[int] = None):
        if width is not None:
            factors = [factors] * width
        self.factors = list(factors)

    def project(self, i: int):
        return self.factors[i]

    def getPField(self):
        return self.factors[0].getPField()

    def getZERO(self):
        return PPRingElement(self, [f.getZERO() for f in self.factors])

    def randomElement(self, randomSource, statDist):
        return PPRingElement(self, [f.randomElement(randomSource, statDist) for f in self.factors])

    def randomElementArray(self, size, randomSource, statDist):
        return PPRingElementArray(self, [f.randomElementArray(size, randomSource, statDist) for f in self.factors])

    def toElement(self, btr: ByteTreeReader):
        if btr.isLeaf() or btr.getRemaining() != len(self.factors):
            raise ArithmFormatException(nat.VMX_EFORMAT, "product ring element of wrong arity")
        return PPRingElement(self, [f.toElement(btr.getNextChild()) for f in self.factors])

    def __eq__(self, o):
        return isinstance(o, PPRing) and o.factors == self.factors

    def __hash__(self):
        return hash(tuple(self.factors))


class PPRingElement:
    def __init__(self, ring: PPRing, comps):
        self.ring = ring
        self.comps = list(comps)

    def getPRing(self):
        return self.ring

    def neg(self):
        return PPRingElement(self.ring, [c.neg() for c in self.comps])

    def add(self, o):
        return PPRingElement(self.ring, [a.add(b) for a, b in zip(self.comps, o.comps)])

    def mulAdd(self, v, b):
        return PPRingElement(self.ring, [a.mulAdd(v, bb) for a, bb in zip(self.comps, b.comps)])

    def toByteTree(self):
        return ByteTreeContainer(*[c.toByteTree() for c in self.comps])

    def equals(self, o):
        return isinstance(o, PPRingElement) and all(a.equals(b) for a, b in zip(self.comps, o.comps))


class PPRingElementArray:
    def __init__(self, ring: PPRing, comps):
        self.ring = ring
        self.comps = list(comps)

    def getPRing(self):
        return self.ring

    def size(self):
        return self.comps[0].size()

    def innerProduct(self, o):
        """Component-wise inner product with a Z_q array (PoSBasicTW.java:863)."""
        if isinstance(o, PPRingElementArray):
            return PPRingElement(self.ring, [a.innerProduct(b) for a, b in zip(self.comps, o.comps)])
        return PPRingElement(self.ring, [a.innerProduct(o) for a in self.comps])

    def free(self):
        for c in self.comps:
            c.free()

    def toByteTree(self):
        return ByteTreeContainer(*[c.toByteTree() for c in self.comps])


class PPGroup(PGroup):
    """Product group: PPGroup(G, k) = G^k, or a product of given factors."""

    def __init__(self, factors, width: Optional[int] = None):
        if width is not None:
            factors = [factors] * width
        self.factors = list(factors)

    def project(self, i: int) -> PGroup:
        return self.factors[i]

    def getWidth(self) -> int:
        return len(self.factors)

    def getPRing(self):
        return PPRing([f.getPRing() for f in self.factors])

    def product(self, *els):
        if len(els) == 1 and not isinstance(els[0], (PGroupElementArray, PPGroupElementArray)):
            els = [els[0]] * len(self.factors)
        if isinstance(els[0], (PGroupElementArray, PPGroupElementArray)):
            return PPGroupElementArray(self, list(els))
        return PPGroupElement(self, list(els))

    def getONE(self):
        return PPGroupElement(self, [f.getONE() for f in self.factors])

    def getg(self):
        return PPGroupElement(self, [f.getg() for f in self.factors])

    def basic(self) -> List[ModPGroup]:
        out = []
        for f in self.factors:
            out += f.basic()
        return out

    def toElement(self, btr: ByteTreeReader):
        if btr.isLeaf() or btr.getRemaining() != len(self.factors):
            raise ArithmFormatException(nat.VMX_EFORMAT, "product element of wrong arity")
        return PPGroupElement(self, [f.toElement(btr.getNextChild()) for f in self.factors])

    def toElementArray(self, size: int, src, check_membership: Optional[bool] = None):
        if isinstance(src, ByteTreeReader):
            if src.isLeaf() or src.getRemaining() != len(self.factors):
                raise ArithmFormatException(nat.VMX_EFORMAT, "product array of wrong arity")
            comps = []
            try:
                for f in self.factors:
                    comps.append(f.toElementArray(size, src.getNextChild(), check_membership=check_membership)
                                 if isinstance(f, ModPGroup) else f.toElementArray(size, src.getNextChild(),
                                                                                    check_membership))
            except Exception:
                for c in comps:
                    c.free()
                raise
            return PPGroupElementArray(self, comps)
        return PPGroupElementArray(self, [f.toElementArray(size, c) for f, c in zip(self.factors, src.comps)])

    def __eq__(self, o):
        return isinstance(o, PPGroup) and o.factors == self.factors

    def __hash__(self):
        return hash(tuple(self.factors))


class PPGroupElement:
    def __init__(self, group: PPGroup, comps):
        self.group = group
        self.comps = list(comps)

    def getPGroup(self):
        return self.group

    def project(self, i: int):
        return self.comps[i]

    def _split(self, e):
        """An exponent from this group's own (product) ring acts component-wise; an exponent from
        any other ring is applied to every component (PPGroupElement.exp in VCR)."""
        if isinstance(e, (PPRingElement, PPRingElementArray)) and e.getPRing() == self.group.getPRing():
            return e.comps
        return [e] * len(self.comps)

    def exp(self, e):
        parts = [c.exp(x) for c, x in zip(self.comps, self._split(e))]
        if isinstance(parts[0], (PGroupElementArray, PPGroupElementArray)):
            return PPGroupElementArray(self.group, parts)
        return PPGroupElement(self.group, parts)

    def mul(self, o):
        return PPGroupElement(self.group, [a.mul(b) for a, b in zip(self.comps, o.comps)])

    def inv(self):
        return PPGroupElement(self.group, [a.inv() for a in self.comps])

    def div(self, o):
        return self.mul(o.inv())

    def expMul(self, e, b):
        return self.exp(e).mul(b)

    def equals(self, o):
        return isinstance(o, PPGroupElement) and len(o.comps) == len(self.comps) and \
            all(a.equals(b) for a, b in zip(self.comps, o.comps))

    __eq__ = equals

    def __hash__(self):
        return hash(tuple(self.comps))

    def toByteTree(self):
        return ByteTreeContainer(*[c.toByteTree() for c in self.comps])


class PPGroupElementArray:
    """Array over a product group, stored column-wise: one device array per factor
    (elgamal/ProtocolElGamalInterfaceRaw.java:53-56)."""

    def __init__(self, group: PPGroup, comps):
        self.group = group
        self.comps = list(comps)

    def getPGroup(self):
        return self.group

    def size(self):
        return self.comps[0].size()

    def project(self, i: int):
        return self.comps[i]

    def basic(self) -> List[PGroupElementArray]:
        out = []
        for c in self.comps:
            out += c.basic()
        return out

    def free(self):
        for c in self.comps:
            c.free()

    def _map(self, fn, *others):
        return PPGroupElementArray(self.group, [fn(c, *[o.comps[i] for o in others])
                                                for i, c in enumerate(self.comps)])

    def mul(self, o):
        return self._map(lambda a, b: a.mul(b), o)

    def inv(self):
        return self._map(lambda a: a.inv())

    def exp(self, e):
        if isinstance(e, (PPRingElement, PPRingElementArray)) and e.getPRing() == self.group.getPRing():
            return PPGroupElementArray(self.group, [c.exp(x) for c, x in zip(self.comps, e.comps)])
        return self._map(lambda a: a.exp(e))

    def expProd(self, e) -> PPGroupElement:
        flat = self.basic()
        res = expProdMany(flat, e)
        it = iter(res)

        def rebuild(arr):
            if isinstance(arr, PPGroupElementArray):
                return PPGroupElement(arr.group, [rebuild(c) for c in arr.comps])
            return next(it)
        return rebuild(self)

    def prod(self):
        return PPGroupElement(self.group, [c.prod() for c in self.comps])

    def permute(self, pi):
        return self._map(lambda a: a.permute(pi))

    def shiftPush(self, el):
        return PPGroupElementArray(self.group, [c.shiftPush(x) for c, x in zip(self.comps, el.comps)])

    def extract(self, keep):
        return self._map(lambda a: a.extract(keep))

    def copyOfRange(self, a, b):
        return self._map(lambda c: c.copyOfRange(a, b))

    def get(self, i):
        return PPGroupElement(self.group, [c.get(i) for c in self.comps])

    def equals(self, o):
        return all(a.equals(b) for a, b in zip(self.comps, o.comps))

    def toByteTree(self):
        return ByteTreeContainer(*[c.toByteTree() for c in self.comps])
