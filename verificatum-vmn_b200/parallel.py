"""Sharding of the hot path over the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch; gloo in the CPU tests).  An
array of global size n lives as contiguous index ranges: rank r owns [r*n/W, (r+1)*n/W).  The
classes below are the `arithm` array classes with the same methods and the same results as on
one GPU, so that `hvzk.PoSBasicTW`, `PoSCBasicTW`, `CCPoSBasicW`, `elgamal.DistrElGamalSessionBasic`
and `mixnet.ShufflerSession` run on them unchanged:

    element-wise ops (exp, mul, inv, add, mulAdd ...)   local, no communication
    expProd / prod / innerProduct / sum / prod          local partial -> all-gather of W single elements
                                                        (W x 384 B) -> multiplied in rank order; the group is
                                                        commutative and exact: bit-identical to one GPU
    equals                                              AND over ranks
    shiftPush                                           one-element halo from the left neighbour
    recLin / prods (Z_q scans)                          local scan + W affine carries
    permute                                             all-to-all of element rows (NCCL) between
                                                        vmx_*_pack_rows / vmx_*_unpack_rows
    random arrays (PRGHeuristic SHA-256)                counter mode: every rank expands its own slice
    toByteTree / to_matrix                              all-gather of the serialised shards (Fiat-Shamir
                                                        hashing is one SHA-256 stream: not shardable)

Single elements (generators, public keys, proof scalars) are replicated: every rank computes the
same O(1) values.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import struct
from typing import List, Optional, Sequence

import numpy as np

from . import _native as nat
from .arithm import (ArithmFormatException, ByteTreeDeviceArray, ECqPGroup, LargeIntegerArray, ModPGroup, Permutation, PField,
                     PFieldElement, PGroupElement, PGroupElementArray, PRingElementArray, _advance_prg, _be, _ptr,
                     _sha256_prg_offset)
from .eio import ByteTreeReader, EIOException


def shard_bounds(n: int, world: int) -> List[int]:
    return [n * r // world for r in range(world + 1)]


class Comm:
    """The process group of one box.  `device` is None for the CPU (gloo + host-emulation) tests."""

    def __init__(self, group=None, device=None, stream_ptr: Optional[int] = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device("cpu") if device is None else torch.device("cuda", device)
        self._stream = None
        if device is not None and stream_ptr:
            self._stream = torch.cuda.ExternalStream(stream_ptr, device=self.device)
        self.bytes_exchanged = 0
        self.collectives = 0

    def on_stream(self):
        """Collectives are ordered with the engine's kernels by making the ctx stream current."""
        if self._stream is None:
            return contextlib.nullcontext()
        return self.torch.cuda.stream(self._stream)

    def allgather_bytes(self, b: bytes) -> List[bytes]:
        """Every rank contributes len(b) bytes (the same length on all ranks)."""
        torch = self.torch
        if self.world == 1:
            return [bytes(b)]
        with self.on_stream():
            src = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(self.device)
            dst = torch.empty(self.world * len(b), dtype=torch.uint8, device=self.device)
            self.dist.all_gather_into_tensor(dst, src, group=self.group)
            raw = dst.cpu().numpy().tobytes()
        self.collectives += 1
        self.bytes_exchanged += len(raw)
        return [raw[i * len(b):(i + 1) * len(b)] for i in range(self.world)]

    def broadcast_bytes(self, b: Optional[bytes], nbytes: int, root: int = 0) -> bytes:
        """`nbytes` bytes of the root rank to every rank (Fiat-Shamir digests: the root alone hashes)."""
        torch = self.torch
        if self.world == 1:
            return bytes(b)
        with self.on_stream():
            if self.rank == root:
                assert b is not None and len(b) == nbytes
                t = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(self.device)
            else:
                t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.dist.broadcast(t, self._global(root), group=self.group)
            out = t.cpu().numpy().tobytes()
        self.collectives += 1
        self.bytes_exchanged += nbytes
        return out

    def allgather_matrix(self, m: np.ndarray, bounds: Sequence[int]) -> np.ndarray:
        """Rows of all ranks, concatenated in rank order (ragged: bounds gives the row ranges)."""
        torch = self.torch
        if self.world == 1:
            return m
        w = m.shape[1]
        counts = [bounds[r + 1] - bounds[r] for r in range(self.world)]
        mx = max(counts)
        if min(counts) == mx and mx:
            # equal shards: the gathered tensor IS the concatenation.  It is read back into one page-locked
            # buffer (DMA) and handed on without a further copy -- these are the serialisations that get hashed.
            with self.on_stream():
                src = torch.from_numpy(np.ascontiguousarray(m).reshape(-1)).to(self.device)
                dst = torch.empty(self.world * mx * w, dtype=torch.uint8, device=self.device)
                self.dist.all_gather_into_tensor(dst, src, group=self.group)
                full = self.host_buffer(self.world * mx * w)
                torch.from_numpy(full).copy_(dst)
            self.collectives += 1
            self.bytes_exchanged += full.nbytes
            return full.reshape(self.world * mx, w)
        with self.on_stream():
            src = torch.zeros((mx, w), dtype=torch.uint8, device=self.device)   # padded to the largest shard
            src[:m.shape[0]] = torch.from_numpy(np.ascontiguousarray(m)).to(self.device)
            dst = torch.empty(self.world * mx * w, dtype=torch.uint8, device=self.device)
            self.dist.all_gather_into_tensor(dst, src.reshape(-1), group=self.group)
            host = dst.cpu().numpy().reshape(self.world, mx, w)
        full = np.concatenate([host[r, :counts[r]] for r in range(self.world)])
        self.collectives += 1
        self.bytes_exchanged += full.nbytes
        return full

    def host_buffer(self, nbytes: int) -> np.ndarray:
        """Host destination of a gathered serialisation: pooled page-locked memory beside a GPU."""
        if self.device.type != "cuda":
            return np.empty(nbytes, dtype=np.uint8)
        from .arithm import _host_buffer
        return _host_buffer(self.device.index or 0, nbytes)

    def all_and(self, flag: bool) -> bool:
        if self.world == 1:
            return bool(flag)
        return all(x == b"\x01" for x in self.allgather_bytes(b"\x01" if flag else b"\x00"))

    def all_max(self, v: int) -> int:
        if self.world == 1:
            return v
        return max(int.from_bytes(x, "big") for x in self.allgather_bytes(int(v).to_bytes(8, "big")))

    def exchange_rows(self, send, send_counts: Sequence[int], recv_counts: Sequence[int]):
        """all-to-all of element rows (int32 [count, limbs] tensors on the engine's device)."""
        torch = self.torch
        recv = torch.empty((int(sum(recv_counts)), send.shape[1]), dtype=send.dtype, device=send.device)
        if self.world == 1:
            recv.copy_(send)
            return recv
        with self.on_stream():
            if self.device.type == "cuda":
                self.dist.all_to_all_single(recv, send, [int(c) for c in recv_counts], [int(c) for c in send_counts],
                                            group=self.group)
            else:  # gloo (CPU tests): pairwise send / recv
                reqs = []
                so = np.concatenate([[0], np.cumsum(send_counts)]).astype(int)
                ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(int)
                recv[ro[self.rank]:ro[self.rank + 1]] = send[so[self.rank]:so[self.rank + 1]]
                for peer in range(self.world):
                    if peer == self.rank:
                        continue
                    if send_counts[peer]:
                        reqs.append(self.dist.isend(send[so[peer]:so[peer + 1]].contiguous(), self._global(peer),
                                                    group=self.group))
                    if recv_counts[peer]:
                        reqs.append(self.dist.irecv(recv[ro[peer]:ro[peer + 1]], self._global(peer), group=self.group))
                for r in reqs:
                    r.wait()
        self.collectives += 1
        self.bytes_exchanged += int(send.numel()) * 4
        return recv

    def _global(self, peer: int) -> int:
        return self.dist.get_global_rank(self.group, peer) if self.group is not None else peer


# ====================================================================== sharded Z_q arrays
class ShardedField(PField):
    def __init__(self, group: "ShardedModPGroup"):
        super().__init__(group)
        self.comm = group.comm

    def _rarr(self, h, size: Optional[int] = None):
        return ShardedRingArray(self, h, size)

    def _range(self, size: int):
        b = shard_bounds(size, self.comm.world)
        return b[self.comm.rank], b[self.comm.rank + 1]

    def randomElementArray(self, size: int, randomSource, statDist: int):
        bits = self.order.bit_length() + statDist
        width = (bits + 7) // 8
        lo, hi = self._range(size)
        off = _sha256_prg_offset(randomSource)
        h = C.c_void_p()
        if off is not None:
            nat.check(nat.load().vmx_rarr_prg_raw_sha256(self.group.ctx, randomSource.seed, len(randomSource.seed),
                                                         off + lo * width, hi - lo, width, bits, C.byref(h)))
            _advance_prg(randomSource, off + size * width)
            return self._rarr(h, size)
        raw = np.frombuffer(randomSource.getBytes(size * width), dtype=np.uint8)[lo * width:hi * width]
        nat.check(nat.load().vmx_rarr_from_raw(self.group.ctx, hi - lo, _ptr(np.ascontiguousarray(raw)), width, bits,
                                               C.byref(h)))
        return self._rarr(h, size)

    def _lia_random(self, size: int, bitLength: int, randomSource) -> LargeIntegerArray:
        lib = nat.load()
        width = (bitLength + 7) // 8
        lo, hi = self._range(size)
        h = C.c_void_p()
        off = _sha256_prg_offset(randomSource)
        if off is not None:  # integers of any width up to 4 residues: reduced mod q on the device
            nat.check(lib.vmx_rarr_prg_raw_sha256(self.group.ctx, randomSource.seed, len(randomSource.seed),
                                                  off + lo * width, hi - lo, width, bitLength, C.byref(h)))
            _advance_prg(randomSource, off + size * width)
            return LargeIntegerArray(self, h, size)
        raw = np.frombuffer(randomSource.getBytes(size * width), dtype=np.uint8)[lo * width:hi * width]
        nat.check(lib.vmx_rarr_from_raw(self.group.ctx, hi - lo, _ptr(np.ascontiguousarray(raw)), width, bitLength,
                                        C.byref(h)))
        return LargeIntegerArray(self, h, size)

    def toElementArray(self, *args):
        lib = nat.load()
        if len(args) == 1 and isinstance(args[0], LargeIntegerArray):
            return args[0]._to_ring(self)
        h = C.c_void_p()
        if len(args) == 1:  # a replicated list of elements: keep this rank's slice
            vals = [int(e.value) for e in args[0]]
            lo, hi = self._range(len(vals))
            m = np.frombuffer(b"".join(_be(v, self.byte_len) for v in vals[lo:hi]), dtype=np.uint8)
            nat.check(lib.vmx_rarr_from_bytes(self.group.ctx, hi - lo, _ptr(m), C.byref(h)))
            return self._rarr(h, len(vals))
        size, src = args
        lo, hi = self._range(size)
        if isinstance(src, ByteTreeReader):
            try:
                stream = src.leaf_stream(size, self.byte_len)
            except EIOException as e:
                raise ArithmFormatException(nat.VMX_EFORMAT, str(e))
            w = 5 + self.byte_len
            ok = True
            try:
                nat.check(lib.vmx_rarr_from_leaves(self.group.ctx, hi - lo, _ptr(stream[lo * w:hi * w]), C.byref(h)))
            except ArithmFormatException:
                ok = False
            if not self.comm.all_and(ok):  # a malformed element in ANY shard rejects the array on every rank
                if ok:
                    lib.vmx_rarr_free(h)
                raise ArithmFormatException(nat.VMX_EFORMAT, "ring element out of range")
            arr = self._rarr(h, size)
            arr._leaves = stream  # every rank parsed the same bytes: no gather needed to hash them
            return arr
        else:
            nat.check(lib.vmx_rarr_fill(self.group.ctx, hi - lo, _be(src.value, self.byte_len), C.byref(h)))
        return self._rarr(h, size)

    unsafeToElementArray = toElementArray


class ShardedRingArray(PRingElementArray):
    _export_into_messages = False   # the serialisation of a sharded array is a gather of all shards (leaves())

    def __init__(self, ring: ShardedField, handle, gsize: int):
        super().__init__(ring, handle)
        self.gsize = int(gsize)
        self.comm = ring.comm
        self.bounds = shard_bounds(self.gsize, self.comm.world)
        self.lo, self.hi = self.bounds[self.comm.rank], self.bounds[self.comm.rank + 1]

    def size(self) -> int:
        return self.gsize

    def local_size(self) -> int:
        return self.hi - self.lo

    def _new(self, h):
        return ShardedRingArray(self.ring, h, self.gsize)

    def _local(self) -> PRingElementArray:
        """A non-owning plain view of the local shard."""
        v = PRingElementArray.__new__(PRingElementArray)
        v.ring, v.h, v._lib = self.ring, self.h, self._lib
        v.free = lambda: None
        return v

    def _gather_scalars(self, x: PFieldElement) -> List[int]:
        return [int.from_bytes(b, "big") for b in self.comm.allgather_bytes(_be(x.value, self.ring.byte_len))]

    # -- reductions
    def innerProduct(self, o) -> PFieldElement:
        loc = PRingElementArray.innerProduct(self, o) if self.local_size() else self.ring.getZERO()
        return PFieldElement(self.ring, sum(self._gather_scalars(loc)))

    def sum(self) -> PFieldElement:
        loc = PRingElementArray.sum(self) if self.local_size() else self.ring.getZERO()
        return PFieldElement(self.ring, sum(self._gather_scalars(loc)))

    def prod(self) -> PFieldElement:
        loc = PRingElementArray.prod(self) if self.local_size() else self.ring.getONE()
        r = 1
        for v in self._gather_scalars(loc):
            r = r * v % self.ring.order
        return PFieldElement(self.ring, r)

    # -- scans: local scan + carries
    def prods(self):
        q = self.ring.order
        y = PRingElementArray.prods(self)
        last = PRingElementArray.get(y, self.local_size() - 1) if self.local_size() else self.ring.getONE()
        totals = self._gather_scalars(last)
        pref = 1
        for v in totals[:self.comm.rank]:
            pref = pref * v % q
        if pref != 1 and self.local_size():
            y2 = y.mul(PFieldElement(self.ring, pref))
            y.free()
            y = y2
        return y

    def recLin(self, e: "ShardedRingArray"):
        """x[i] = x[i-1]*e[i] + b[i]: on a shard x = xloc + X_prev * prods(e_loc), X_prev = x at the
        last index of the left neighbour (an affine carry per rank)."""
        q = self.ring.order
        n = self.local_size()
        if n:
            xloc, dloc = PRingElementArray.recLin(self, e)
            yloc = PRingElementArray.prods(e)
            a_tot = PRingElementArray.get(yloc, n - 1)
        else:
            xloc, dloc, yloc, a_tot = self._new(self._empty_handle()), self.ring.getZERO(), None, self.ring.getONE()
        A = self._gather_scalars(a_tot)
        B = self._gather_scalars(dloc)
        X = 0
        x_prev = 0
        for r in range(self.comm.world):
            if r == self.comm.rank:
                x_prev = X
            X = (X * A[r] + B[r]) % q
        if n and x_prev:
            x = yloc.mulAdd(PFieldElement(self.ring, x_prev), xloc)
            xloc.free()
        else:
            x = xloc
        if yloc is not None:
            yloc.free()
        return x, PFieldElement(self.ring, X)

    def _empty_handle(self):
        h = C.c_void_p()
        nat.check(self._lib.vmx_rarr_from_bytes(self.ring.group.ctx, 0, None, C.byref(h)))
        return h

    # -- data movement
    def permute(self, pi: Permutation):
        return _sharded_permute(self, pi, ring=True)

    def shiftPush(self, el: PFieldElement):
        last = PRingElementArray.get(self, self.local_size() - 1) if self.local_size() else self.ring.getZERO()
        lasts = self._gather_scalars(last)
        first = el.value
        for r in range(self.comm.rank):
            if self.bounds[r + 1] > self.bounds[r]:
                first = lasts[r]
        if not self.local_size():
            return self._new(self._empty_handle())
        return PRingElementArray.shiftPush(self, PFieldElement(self.ring, first))

    def copyOfRange(self, a: int, b: int):
        if (a, b) == (0, self.gsize):
            return PRingElementArray.copyOfRange(self, 0, self.local_size())
        return _sharded_select(self, np.arange(a, b, dtype=np.int64), ring=True)

    def extract(self, keep: Sequence[bool]):
        return _sharded_select(self, _kept_indices(keep, self.gsize), ring=True)

    def get(self, i: int) -> PFieldElement:
        own = self.lo <= i < self.hi
        v = PRingElementArray.get(self, i - self.lo) if own else self.ring.getZERO()
        owner = next(r for r in range(self.comm.world) if self.bounds[r] <= i < self.bounds[r + 1])
        return PFieldElement(self.ring, self._gather_scalars(v)[owner])

    def equals(self, o) -> bool:
        return self.comm.all_and(PRingElementArray.equals(self, o) if self.local_size() else True)

    def bitLength(self) -> int:
        return self.comm.all_max(PRingElementArray.bitLength(self) if self.local_size() else 0)

    # -- I/O
    def leaves(self) -> np.ndarray:
        """Serialisation of the WHOLE array (all shards, gathered to every rank), cached."""
        if self._leaves is None:
            w = 5 + self.ring.byte_len
            m = self.comm.host_buffer(self.local_size() * w).reshape(self.local_size(), w)
            nat.check(self._lib.vmx_rarr_to_leaves(self.h, _ptr(m)))
            self._leaves = np.ascontiguousarray(self.comm.allgather_matrix(m, self.bounds)).reshape(-1)
        return self._leaves


# ====================================================================== sharded group arrays
class _ShardedGroupMixin:
    """What a sharded group adds to ModPGroup / ECqPGroup: array wrappers, index ranges, combination of the
    ranks' partial products."""

    def _garr(self, h, size: Optional[int] = None):
        return ShardedGroupArray(self, h, size)

    def _range(self, size: int):
        b = shard_bounds(size, self.comm.world)
        return b[self.comm.rank], b[self.comm.rank + 1]

    def _combine_partials(self, parts: List[PGroupElement]) -> List[PGroupElement]:
        """Every rank holds one partial product per component: all-gather (W x k elements) and
        multiply in rank order on the device."""
        W = self.comm.world
        if W == 1:
            return parts
        k = len(parts)
        raw = self.comm.allgather_bytes(b"".join(_be(x.value, self.elem_bytes) for x in parts))
        out = []
        for c in range(k):
            m = np.frombuffer(b"".join(raw[r][c * self.elem_bytes:(c + 1) * self.elem_bytes] for r in range(W)),
                              dtype=np.uint8)
            h = C.c_void_p()
            nat.check(self._lib.vmx_garr_from_bytes(self.ctx, W, _ptr(m), 0, C.byref(h)))
            buf = np.empty(self.elem_bytes, dtype=np.uint8)
            try:
                nat.check(self._lib.vmx_prod(h, _ptr(buf)))
            finally:
                self._lib.vmx_garr_free(h)
            out.append(PGroupElement(self, int.from_bytes(buf.tobytes(), "big")))
        return out


class ShardedModPGroup(_ShardedGroupMixin, ModPGroup):
    def __init__(self, p: int, q: int, g: int, comm_factory, device: int = 0):
        ModPGroup.__init__(self, p, q, g, device=device)
        self.comm = comm_factory(self)
        self.pRing = ShardedField(self)

    def toElementArray(self, *args, check_membership: Optional[bool] = None):
        lib = self._lib
        if check_membership is None:
            check_membership = self.membership_check
        h = C.c_void_p()
        if len(args) == 1:
            vals = [e.value for e in args[0]]
            lo, hi = self._range(len(vals))
            m = np.frombuffer(b"".join(_be(v, self.elem_bytes) for v in vals[lo:hi]), dtype=np.uint8)
            nat.check(lib.vmx_garr_from_bytes(self.ctx, hi - lo, _ptr(m), 0, C.byref(h)))
            return self._garr(h, len(vals))
        size, src = args
        lo, hi = self._range(size)
        if isinstance(src, (ByteTreeReader, np.ndarray)):
            stream = None
            if isinstance(src, ByteTreeReader):
                try:
                    m = src.leaf_matrix(size, self.elem_bytes)
                    stream = src.leaf_stream(size, self.elem_bytes)
                except EIOException as e:
                    raise ArithmFormatException(nat.VMX_EFORMAT, str(e))
            else:
                m = src.reshape(size, self.elem_bytes)
            m = np.ascontiguousarray(m[lo:hi])
            ok = True
            try:
                nat.check(lib.vmx_garr_from_bytes(self.ctx, hi - lo, _ptr(m), 1 if check_membership else 0, C.byref(h)))
            except ArithmFormatException:
                ok = False
            if not self.comm.all_and(ok):
                if ok:
                    lib.vmx_garr_free(h)
                raise ArithmFormatException(nat.VMX_EFORMAT, "group element out of range or not in the subgroup")
            arr = self._garr(h, size)
            if stream is not None:
                arr._leaves = stream  # every rank parsed the same bytes: no gather needed to hash them
            return arr
        else:
            nat.check(lib.vmx_garr_fill(self.ctx, hi - lo, _be(src.value, self.elem_bytes), C.byref(h)))
        return self._garr(h, size)

    def randomElementArray(self, size: int, randomSource, statDist: int):
        bits = self.p.bit_length() + statDist
        width = (bits + 7) // 8
        lo, hi = self._range(size)
        h = C.c_void_p()
        off = _sha256_prg_offset(randomSource)
        if off is not None:
            nat.check(self._lib.vmx_garr_prg_sha256(self.ctx, randomSource.seed, len(randomSource.seed),
                                                    off + lo * width, hi - lo, width, bits, C.byref(h)))
            _advance_prg(randomSource, off + size * width)
            return self._garr(h, size)
        raw = np.frombuffer(randomSource.getBytes(size * width), dtype=np.uint8)[lo * width:hi * width]
        nat.check(self._lib.vmx_garr_from_raw(self.ctx, hi - lo, _ptr(np.ascontiguousarray(raw)), width, bits,
                                              C.byref(h)))
        return self._garr(h, size)


class ShardedECqPGroup(_ShardedGroupMixin, ECqPGroup):
    """ECqPGroup with arrays sharded over the ranks (BASELINE.json config 5: P-256 at 1/2/4/8 GPUs)."""

    def __init__(self, name: str, comm_factory, device: int = 0):
        ECqPGroup.__init__(self, name, device=device)
        self.comm = comm_factory(self)
        self.pRing = ShardedField(self)

    def toElementArray(self, *args, check_membership: Optional[bool] = None):
        lib = self._lib
        h = C.c_void_p()
        cb = self.coord_bytes
        if len(args) == 1:
            vals = [e.value for e in args[0]]
            lo, hi = self._range(len(vals))
            m = np.frombuffer(b"".join(_be(v, self.elem_bytes) for v in vals[lo:hi]), dtype=np.uint8)
            nat.check(lib.vmx_garr_from_bytes(self.ctx, hi - lo, _ptr(m), 0, C.byref(h)))
            return self._garr(h, len(vals))
        size, src = args
        lo, hi = self._range(size)
        if isinstance(src, (ByteTreeReader, np.ndarray)):
            stream = None
            if isinstance(src, ByteTreeReader):
                try:
                    stream = src.point_array_stream(size, cb)
                except EIOException as e:
                    raise ArithmFormatException(nat.VMX_EFORMAT, str(e))
                run = size * (5 + cb)
                hdr = np.frombuffer(struct.pack(">BI", 0, size), dtype=np.uint8)
                leaf = np.frombuffer(struct.pack(">BI", 1, cb), dtype=np.uint8)
                xs = stream[5:5 + run].reshape(size, 5 + cb)
                ys = stream[10 + run:].reshape(size, 5 + cb)
                if not (np.array_equal(stream[:5], hdr) and np.array_equal(stream[5 + run:10 + run], hdr) and
                        (xs[:, :5] == leaf).all() and (ys[:, :5] == leaf).all()):
                    raise ArithmFormatException(nat.VMX_EFORMAT, "point array: malformed coordinate nodes")
                m = np.concatenate([xs[lo:hi, 5:], ys[lo:hi, 5:]], axis=1)
            else:
                m = src.reshape(size, self.elem_bytes)[lo:hi]
            m = np.ascontiguousarray(m)
            ok = True
            try:
                nat.check(lib.vmx_garr_from_bytes(self.ctx, hi - lo, _ptr(m), 1, C.byref(h)))
            except ArithmFormatException:
                ok = False
            if not self.comm.all_and(ok):
                if ok:
                    lib.vmx_garr_free(h)
                raise ArithmFormatException(nat.VMX_EFORMAT, "not a point of the curve")
            arr = self._garr(h, size)
            if stream is not None:
                arr._leaves = stream
            return arr
        nat.check(lib.vmx_garr_fill(self.ctx, hi - lo, _be(src.value, self.elem_bytes), C.byref(h)))
        return self._garr(h, size)

    def randomElementArray(self, size: int, randomSource, statDist: int):
        """Rejection sampling makes the position of point i in the stream depend on all earlier candidates:
        every rank derives the whole array (one square-root test per candidate) and keeps its range."""
        full = self._random_full_handle(size, randomSource, statDist)
        lo, hi = self._range(size)
        h = C.c_void_p()
        try:
            nat.check(self._lib.vmx_slice(full, lo, hi, C.byref(h)))
        finally:
            self._lib.vmx_garr_free(full)
        return self._garr(h, size)


class ShardedGroupArray(PGroupElementArray):
    _export_into_messages = False

    def __init__(self, group: ShardedModPGroup, handle, gsize: int):
        super().__init__(group, handle)
        self.gsize = int(gsize)
        self.comm = group.comm
        self.bounds = shard_bounds(self.gsize, self.comm.world)
        self.lo, self.hi = self.bounds[self.comm.rank], self.bounds[self.comm.rank + 1]

    def size(self) -> int:
        return self.gsize

    def local_size(self) -> int:
        return self.hi - self.lo

    def _new(self, h):
        return ShardedGroupArray(self.group, h, self.gsize)

    def _gather_elems(self, x: PGroupElement) -> List[int]:
        return [int.from_bytes(b, "big") for b in self.comm.allgather_bytes(_be(x.value, self.group.elem_bytes))]

    def _empty_handle(self):
        h = C.c_void_p()
        nat.check(self._lib.vmx_garr_from_bytes(self.group.ctx, 0, None, 0, C.byref(h)))
        return h

    def permute(self, pi: Permutation):
        return _sharded_permute(self, pi, ring=False)

    def shiftPush(self, el: PGroupElement):
        last = PGroupElementArray.get(self, self.local_size() - 1) if self.local_size() else self.group.getONE()
        lasts = self._gather_elems(last)
        first = el.value
        for r in range(self.comm.rank):
            if self.bounds[r + 1] > self.bounds[r]:
                first = lasts[r]
        if not self.local_size():
            return self._new(self._empty_handle())
        return PGroupElementArray.shiftPush(self, PGroupElement(self.group, first))

    def extract(self, keep: Sequence[bool]):
        """mixnet/PermutationCommitment.java:462-468 (`commitment.extract(keepList)`): the kept elements in order;
        the result is sharded over its own (smaller) index range, so kept elements move between ranks."""
        return _sharded_select(self, _kept_indices(keep, self.gsize), ring=False)

    def copyOfRange(self, a: int, b: int):
        if (a, b) == (0, self.gsize):
            return PGroupElementArray.copyOfRange(self, 0, self.local_size())
        return _sharded_select(self, np.arange(a, b, dtype=np.int64), ring=False)

    def get(self, i: int) -> PGroupElement:
        own = self.lo <= i < self.hi
        v = PGroupElementArray.get(self, i - self.lo) if own else self.group.getONE()
        owner = next(r for r in range(self.comm.world) if self.bounds[r] <= i < self.bounds[r + 1])
        return PGroupElement(self.group, self._gather_elems(v)[owner])

    def equals(self, o) -> bool:
        return self.comm.all_and(PGroupElementArray.equals(self, o) if self.local_size() else True)

    def leaves(self) -> np.ndarray:
        """Serialisation of the WHOLE array (all shards, gathered to every rank), cached."""
        if self._leaves is None and self.group.is_curve:
            cb, nloc = self.group.coord_bytes, self.local_size()
            buf = self.comm.host_buffer(self.group._leaves_bytes(nloc))
            nat.check(self._lib.vmx_garr_to_leaves(self.h, _ptr(buf)))
            run = nloc * (5 + cb)
            gx = self.comm.allgather_matrix(buf[5:5 + run].reshape(nloc, 5 + cb), self.bounds)
            gy = self.comm.allgather_matrix(buf[10 + run:].reshape(nloc, 5 + cb), self.bounds)
            hdr = np.frombuffer(struct.pack(">BI", 0, self.gsize), dtype=np.uint8)
            self._leaves = np.concatenate([hdr, gx.reshape(-1), hdr, gy.reshape(-1)])
        if self._leaves is None:
            w = 5 + self.group.elem_bytes
            m = self.comm.host_buffer(self.local_size() * w).reshape(self.local_size(), w)
            nat.check(self._lib.vmx_garr_to_leaves(self.h, _ptr(m)))
            self._leaves = np.ascontiguousarray(self.comm.allgather_matrix(m, self.bounds)).reshape(-1)
        return self._leaves

    def to_matrix(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if not self.group.is_curve:
            return PGroupElementArray.to_matrix(self, out)
        cb, n = self.group.coord_bytes, self.gsize
        lv = self.leaves()
        run = n * (5 + cb)
        m = np.concatenate([lv[5:5 + run].reshape(n, 5 + cb)[:, 5:], lv[10 + run:].reshape(n, 5 + cb)[:, 5:]], axis=1)
        if out is not None:
            out[:] = m
            return out
        return m


def _sharded_permute(arr, pi: Permutation, ring: bool):
    """out[pi[i]] = a[i] over GLOBAL indices.  Local elements are packed element-major in the order
    of their destination rank, exchanged with one all-to-all, and scattered to their local
    destination slots; the (replicated) permutation table gives every rank both lists."""
    comm = arr.comm
    lib = arr._lib
    torch = comm.torch
    ctx = arr.ring.group.ctx if ring else arr.group.ctx
    nl = int(lib.vmx_ctx_ring_row_bytes(ctx) if ring else lib.vmx_ctx_row_bytes(ctx)) // 4
    tbl = pi.table
    if tbl.shape[0] != arr.gsize:
        raise nat.VmxError(nat.VMX_ESIZE, "permutation of the wrong size")
    b = np.asarray(arr.bounds, dtype=np.int64)
    lo, hi = arr.lo, arr.hi
    # the routing plan depends on the table and the shard bounds only: one permutation moves several arrays
    # (both ciphertext components, the commitment, the batching vector), so it is kept on the Permutation
    plans = pi.__dict__.setdefault("_shard_plans", {})
    key = (arr.gsize, comm.rank, comm.world)
    if key not in plans:
        dest = tbl[lo:hi]
        dest_rank = (np.searchsorted(b[1:], dest, side="right")).astype(np.int32)
        order = np.argsort(dest_rank, kind="stable").astype(np.uint32)
        send_counts = np.bincount(dest_rank, minlength=comm.world).astype(np.int64)
        mine = np.flatnonzero((tbl >= lo) & (tbl < hi))          # global sources landing here, in source order
        src_rank = np.searchsorted(b[1:], mine, side="right")
        recv_counts = np.bincount(src_rank, minlength=comm.world).astype(np.int64)
        dst_idx = np.ascontiguousarray((tbl[mine].astype(np.int64) - lo).astype(np.uint32))
        plans[key] = (order, send_counts, recv_counts, dst_idx)
    order, send_counts, recv_counts, dst_idx = plans[key]
    pack = lib.vmx_rarr_pack_rows if ring else lib.vmx_garr_pack_rows
    unpack = lib.vmx_rarr_unpack_rows if ring else lib.vmx_garr_unpack_rows
    with comm.on_stream():
        send = torch.empty((hi - lo, nl), dtype=torch.int32, device=comm.device)
        if hi - lo:
            nat.check(pack(arr.h, _ptr(np.ascontiguousarray(order)), hi - lo, C.c_void_p(send.data_ptr())))
        recv = comm.exchange_rows(send, send_counts, recv_counts)
        h = C.c_void_p()
        nat.check(unpack(ctx, hi - lo, C.c_void_p(recv.data_ptr()) if hi - lo else None,
                         _ptr(dst_idx) if hi - lo else None, hi - lo, C.byref(h)))
        if comm.device.type == "cuda":
            # the rows must stay alive until the unpack kernel has read them
            recv.record_stream(torch.cuda.current_stream())
            send.record_stream(torch.cuda.current_stream())
    return arr._new(h)


def _kept_indices(keep, size: int) -> np.ndarray:
    flags = np.asarray(keep, dtype=bool)
    if flags.shape[0] != size:
        raise nat.VmxError(nat.VMX_ESIZE, "keep list of the wrong size")
    return np.flatnonzero(flags).astype(np.int64)


def _sharded_select(arr, src: np.ndarray, ring: bool):
    """out[j] = a[src[j]] for an increasing list of GLOBAL source indices (extract, copyOfRange): stream
    compaction across shards.  `src` is replicated host data (a keep list, a range), so every rank derives the
    whole routing from it -- the "local compaction + exclusive scan of counts" of SURVEY.md section 8e without a
    collective for the counts: rank r packs its selected elements in output order (the destination rank is
    monotone in j), one all-to-all of element rows moves them, and the receiver scatters them to their slots of
    the new, smaller index range."""
    comm = arr.comm
    lib = arr._lib
    torch = comm.torch
    ctx = arr.ring.group.ctx if ring else arr.group.ctx
    nl = int(lib.vmx_ctx_ring_row_bytes(ctx) if ring else lib.vmx_ctx_row_bytes(ctx)) // 4
    m = int(src.shape[0])
    if m and (src[0] < 0 or src[-1] >= arr.gsize or (m > 1 and (np.diff(src) <= 0).any())):
        raise nat.VmxError(nat.VMX_EARG, "selection out of range or not increasing")
    ob = np.asarray(arr.bounds, dtype=np.int64)
    nb = np.asarray(shard_bounds(m, comm.world), dtype=np.int64)
    lo, hi = arr.lo, arr.hi
    nlo, nhi = int(nb[comm.rank]), int(nb[comm.rank + 1])
    j0, j1 = np.searchsorted(src, lo, side="left"), np.searchsorted(src, hi, side="left")   # outputs sourced here
    mine_j = np.arange(j0, j1, dtype=np.int64)
    order = np.ascontiguousarray((src[j0:j1] - lo).astype(np.uint32))                       # local rows, output order
    dest_rank = np.searchsorted(nb[1:], mine_j, side="right")
    send_counts = np.bincount(dest_rank, minlength=comm.world).astype(np.int64)
    src_rank = np.searchsorted(ob[1:], src[nlo:nhi], side="right")                          # monotone in j as well
    recv_counts = np.bincount(src_rank, minlength=comm.world).astype(np.int64)
    nloc = nhi - nlo
    dst_idx = np.arange(nloc, dtype=np.uint32)    # rows arrive by source rank, i.e. already in output order
    pack = lib.vmx_rarr_pack_rows if ring else lib.vmx_garr_pack_rows
    unpack = lib.vmx_rarr_unpack_rows if ring else lib.vmx_garr_unpack_rows
    with comm.on_stream():
        send = torch.empty((int(order.shape[0]), nl), dtype=torch.int32, device=comm.device)
        if order.shape[0]:
            nat.check(pack(arr.h, _ptr(order), int(order.shape[0]), C.c_void_p(send.data_ptr())))
        recv = comm.exchange_rows(send, send_counts, recv_counts)
        h = C.c_void_p()
        nat.check(unpack(ctx, nloc, C.c_void_p(recv.data_ptr()) if nloc else None,
                         _ptr(dst_idx) if nloc else None, nloc, C.byref(h)))
        if comm.device.type == "cuda":
            recv.record_stream(torch.cuda.current_stream())
            send.record_stream(torch.cuda.current_stream())
    return ShardedRingArray(arr.ring, h, m) if ring else ShardedGroupArray(arr.group, h, m)


def make_curve_group(name: str, device: Optional[int], group=None) -> ShardedECqPGroup:
    """An ECqPGroup whose arrays are sharded over the ranks of `group` (default: the world)."""
    def factory(G):
        return Comm(group, device, int(G._lib.vmx_ctx_stream(G.ctx) or 0) if device is not None else None)
    return ShardedECqPGroup(name, factory, device=0 if device is None else device)


def make_group(p: int, q: int, g: int, device: Optional[int], group=None) -> ShardedModPGroup:
    """A ModPGroup whose arrays are sharded over the ranks of `group` (default: the world)."""
    def factory(G):
        return Comm(group, device, int(G._lib.vmx_ctx_stream(G.ctx) or 0) if device is not None else None)
    return ShardedModPGroup(p, q, g, factory, device=0 if device is None else device)
