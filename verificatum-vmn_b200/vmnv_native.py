"""The native universal verifier (csrc/vmnv_native.cpp -> libvmnv.so, C ABI in include/vmnv.h) behind the interface
of `vmnv.MixNetElGamalVerifyFiatShamirSession`: the same report, the same `VerificationError` where the reference
stops with an error (mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668).

The pipeline itself -- byte-tree walking, Fiat-Shamir hashing beside the GPU, the order of the engine calls -- is
C++ over include/vmx.h; this module only marshals the protocol parameters and the files of the proof directory
(zero-copy: bytes, bytearray, numpy arrays and HostBytes are passed by address).  ModPGroup and ECqPGroup, any width.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from . import _native as nat
from .mixnet import SessionParams
from .vmnv import VerificationError

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libvmnv.so")


class _Params(C.Structure):
    _fields_ = [("kind", C.c_int), ("p_be", C.c_char_p), ("q_be", C.c_char_p), ("g_be", C.c_char_p),
                ("a_be", C.c_char_p), ("b_be", C.c_char_p), ("gy_be", C.c_char_p), ("nbytes", C.c_size_t),
                ("device", C.c_int), ("k", C.c_int), ("threshold", C.c_int), ("vbitlenro", C.c_int),
                ("ebitlenro", C.c_int), ("rbitlen", C.c_int), ("version", C.c_char_p), ("sid", C.c_char_p),
                ("pgroup_string", C.c_char_p), ("expected_auxsid", C.c_char_p), ("expected_width", C.c_int),
                ("expected_type", C.c_char_p), ("nodec", C.c_int), ("noposc", C.c_int), ("noccpos", C.c_int)]


class _File(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("size", C.c_size_t)]


class _Report(C.Structure):
    _fields_ = [("accepted", C.c_int), ("fail_stop", C.c_int), ("type", C.c_int), ("n_shuffles", C.c_int),
                ("shuffles", C.c_int * 64), ("poscs", C.c_int * 64), ("valid_proofs", C.c_int),
                ("enough_valid_proofs", C.c_int), ("decryption", C.c_int), ("plaintexts", C.c_int),
                ("hashed_bytes", C.c_uint64), ("launches", C.c_uint64), ("error", C.c_char * 400),
                ("test_vectors", C.c_char * 16384)]


_lib = None


def load():
    """libvmnv.so bound to the engine library this process uses (`_native.load()`: the CUDA build, or the
    host-emulation build when the CPU tests set VMX_LIBRARY_PATH)."""
    global _lib
    if _lib is None:
        path = os.environ.get("VMNV_LIBRARY_PATH", DEFAULT_LIB)
        if not os.path.exists(path):
            raise nat.VmxError(nat.VMX_EARG, "native verifier %s not built (python -c 'import __graft_entry__ as g; g.build()')" % path)
        nat.load()
        lib = C.CDLL(path)
        lib.vmxv_bind.argtypes = [C.c_char_p]
        lib.vmxv_verify.argtypes = [C.POINTER(_Params), C.POINTER(_File), C.c_size_t, C.POINTER(_Report)]
        engine = os.environ.get("VMX_LIBRARY_PATH", nat.DEFAULT_LIB)
        if lib.vmxv_bind(engine.encode()) != 0:
            raise nat.VmxError(nat.VMX_EARG, "libvmnv.so could not bind the engine library %s" % engine)
        _lib = lib
    return _lib


def _address(data):
    """(address, size, keep-alive) of a bytes-like value without copying it."""
    if isinstance(data, bytes):
        return C.cast(C.c_char_p(data), C.c_void_p).value, len(data), data
    a = np.frombuffer(data, dtype=np.uint8)
    return a.ctypes.data, a.size, (a, data)


class MixNetElGamalVerifyFiatShamirSessionNative:
    """Drop-in for vmnv.MixNetElGamalVerifyFiatShamirSession (same options, same report)."""

    def __init__(self, pGroup, params: SessionParams, k: int, threshold: int, expectedAuxsid: Optional[str] = None,
                 expectedWidth: Optional[int] = None, expectedType: Optional[str] = None, dec: bool = True,
                 posc: bool = True, ccpos: bool = True):
        if params.rohash != "SHA-256" or params.prghash != "SHA-256":
            raise NotImplementedError("the native verifier hashes with SHA-256")
        self.pGroup, self.params, self.k, self.threshold = pGroup, params, k, threshold
        self.expectedAuxsid, self.expectedWidth, self.expectedType = expectedAuxsid, expectedWidth, expectedType
        self.dec, self.posc, self.ccpos = dec, posc, ccpos
        self.report: Dict[str, object] = {}

    def verify(self, nizkp) -> Dict[str, object]:
        lib = load()
        G, p = self.pGroup, self.params
        nbytes = (G.p.bit_length() + 7) // 8
        be = lambda v: int(v).to_bytes(nbytes, "big")
        if getattr(G, "is_curve", False):
            group = (1, be(G.p), be(G.q), be(G.gx), be(G.a), be(G.b), be(G.gy))
        else:
            group = (0, be(G.p), be(G.q), be(G.g), None, None, None)
        P = _Params(*group, nbytes, getattr(G, "device", 0), self.k, self.threshold, p.vbitlenro,
                    p.ebitlenro, p.rbitlen, p.version.encode(), p.sid.encode(), p.pGroupString.encode(),
                    (self.expectedAuxsid or "").encode(), int(self.expectedWidth or 0),
                    (self.expectedType or "").encode(), int(not self.dec), int(not self.posc), int(not self.ccpos))
        names = sorted(nizkp)
        files = (_File * len(names))()
        keep = []
        for f, name in zip(files, names):
            addr, size, alive = _address(nizkp[name])
            keep.append(alive)
            f.name, f.data, f.size = name.encode(), addr, size
        R = _Report()
        rc = lib.vmxv_verify(C.byref(P), files, len(names), C.byref(R))
        if rc != 0:
            raise nat.VmxError(nat.VMX_ECUDA, "native verifier: %s" % R.error.decode("utf-8", "replace"))
        tri = lambda v: None if v < 0 else bool(v)
        span = min(R.n_shuffles, 64)
        rep = {"type": ("mixing", "shuffling", "decryption")[R.type],
               "shuffles": {l + 1: R.shuffles[l] > 0 for l in range(span) if R.shuffles[l]},
               "poscs": {l + 1: R.poscs[l] > 0 for l in range(span) if R.poscs[l]},
               "decryption": tri(R.decryption), "hashed_bytes": int(R.hashed_bytes), "launches": int(R.launches),
               "validProofs": int(R.valid_proofs), "enoughValidProofs": bool(R.enough_valid_proofs)}
        rep["vectors"] = []
        for ln in R.test_vectors.decode("utf-8", "replace").splitlines():
            head, _, value = ln.partition("=")
            name, _, party = head.partition("@")
            rep["vectors"].append((name, int(party) or None, value))
        self.report = rep
        if R.fail_stop:
            raise VerificationError(R.error.decode("utf-8", "replace"))
        if R.plaintexts >= 0:
            rep["plaintexts"] = bool(R.plaintexts)
        rep["accepted"] = bool(R.accepted)
        return rep
