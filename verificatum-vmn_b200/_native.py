"""ctypes binding of the C ABI in include/vmx.h (the drop-in boundary).

This is the Python analogue of the JNI/FFM stub a VCR maintainer would add:
it declares exactly the symbols of include/vmx.h and nothing else.  The shared library is the
in-tree CUDA build `verificatum-vmn_b200/libvmx.so` (built by `__graft_entry__.build()`).
There is NO fallback: if the library is missing, or no CUDA device is present, every call
fails loudly (`VmxError`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libvmx.so")

VMX_OK, VMX_EFORMAT, VMX_ESIZE, VMX_ENOMEM, VMX_ECUDA, VMX_EARG = range(6)


class VmxError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__("vmx status %d: %s" % (status, msg))
        self.status = status


class ArithmFormatError(VmxError):
    """VMX_EFORMAT: the engine-side analogue of com.verificatum.arithm.ArithmFormatException."""


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_U8 = C.c_char_p  # borrowed host bytes
_SZ = C.c_size_t

# name -> (restype, argtypes); mirrors include/vmx.h one to one
SIGNATURES = {
    "vmx_last_error": (C.c_char_p, []),
    "vmx_version": (C.c_int, []),
    "vmx_ctx_create_modp": (C.c_int, [_U8, _U8, _U8, _SZ, C.c_int, _PP]),
    "vmx_ctx_create_ecq": (C.c_int, [_U8, _U8, _U8, _U8, _U8, _U8, _SZ, C.c_int, _PP]),
    "vmx_ctx_destroy": (None, [_P]),
    "vmx_host_alloc": (C.c_int, [C.c_int, _SZ, _PP]),
    "vmx_host_free": (None, [_P]),
    "vmx_host_pool_bytes": (_SZ, []),
    "vmx_ctx_elem_bytes": (_SZ, [_P]),
    "vmx_ctx_ring_bytes": (_SZ, [_P]),
    "vmx_ctx_sync": (C.c_int, [_P]),
    "vmx_ctx_stream": (_P, [_P]),
    "vmx_ctx_set_fixed_window": (C.c_int, [_P, C.c_int]),
    "vmx_ctx_set_tuning": (C.c_int, [_P, C.c_char_p, C.c_longlong]),
    "vmx_leaves_uniform": (C.c_int, [_P, _SZ, _SZ]),
    "vmx_garr_from_bytes": (C.c_int, [_P, _SZ, _P, C.c_int, _PP]),
    "vmx_garr_from_raw": (C.c_int, [_P, _SZ, _P, _SZ, C.c_uint, _PP]),
    "vmx_garr_prg_sha256": (C.c_int, [_P, _U8, _SZ, C.c_uint64, _SZ, _SZ, C.c_uint, _PP]),
    "vmx_ctx_prg_consumed": (C.c_uint64, [_P]),
    "vmx_garr_from_candidates": (C.c_int, [_P, _SZ, _P, _SZ, C.c_uint, _SZ, _PP, C.POINTER(C.c_size_t)]),
    "vmx_rarr_prg_raw_sha256": (C.c_int, [_P, _U8, _SZ, C.c_uint64, _SZ, _SZ, C.c_uint, _PP]),
    "vmx_prg_bytes_sha256": (C.c_int, [_P, _U8, _SZ, C.c_uint64, _SZ, _P]),
    "vmx_permutation_prg_sha256": (C.c_int, [_P, _U8, _SZ, C.c_uint64, _SZ, _SZ, C.c_uint, _P, C.POINTER(C.c_int)]),
    "vmx_permutation_from_raw": (C.c_int, [_P, _SZ, _P, _SZ, C.c_uint, _P, C.POINTER(C.c_int)]),
    "vmx_garr_to_bytes": (C.c_int, [_P, _P]),
    "vmx_garr_from_leaves": (C.c_int, [_P, _SZ, _P, C.c_int, _PP]),
    "vmx_garr_to_leaves": (C.c_int, [_P, _P]),
    "vmx_rarr_from_leaves": (C.c_int, [_P, _SZ, _P, _PP]),
    "vmx_rarr_to_leaves": (C.c_int, [_P, _P]),
    "vmx_garr_fill": (C.c_int, [_P, _SZ, _U8, _PP]),
    "vmx_garr_free": (None, [_P]),
    "vmx_garr_size": (_SZ, [_P]),
    "vmx_exp_fixed": (C.c_int, [_P, _U8, _P, _PP]),
    "vmx_elem_exp": (C.c_int, [_P, _U8, _U8, _P]),
    "vmx_elem_inv": (C.c_int, [_P, _U8, _P]),
    "vmx_fixed_precompute": (C.c_int, [_P, _U8, _SZ]),
    "vmx_ctx_row_bytes": (_SZ, [_P]),
    "vmx_ctx_ring_row_bytes": (_SZ, [_P]),
    "vmx_garr_pack_rows": (C.c_int, [_P, _P, _SZ, _P]),
    "vmx_garr_unpack_rows": (C.c_int, [_P, _SZ, _P, _P, _SZ, _PP]),
    "vmx_rarr_pack_rows": (C.c_int, [_P, _P, _SZ, _P]),
    "vmx_rarr_unpack_rows": (C.c_int, [_P, _SZ, _P, _P, _SZ, _PP]),
    "vmx_exp_var": (C.c_int, [_P, _P, _PP]),
    "vmx_exp_scalar": (C.c_int, [_P, _U8, _PP]),
    "vmx_exp_scalar_var": (C.c_int, [_P, _U8, _P, _P, _PP]),
    "vmx_expprod": (C.c_int, [_PP, _SZ, _P, _P]),
    "vmx_expprod_cols": (C.c_int, [_PP, _SZ, C.POINTER(C.c_int64), _PP]),
    "vmx_mul": (C.c_int, [_P, _P, _PP]),
    "vmx_inv": (C.c_int, [_P, _PP]),
    "vmx_prod": (C.c_int, [_P, _P]),
    "vmx_permute": (C.c_int, [_P, _P, _PP]),
    "vmx_shift_push": (C.c_int, [_P, _U8, _PP]),
    "vmx_extract": (C.c_int, [_P, _U8, _PP]),
    "vmx_slice": (C.c_int, [_P, _SZ, _SZ, _PP]),
    "vmx_equals": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "vmx_get": (C.c_int, [_P, _SZ, _P]),
    "vmx_rarr_from_bytes": (C.c_int, [_P, _SZ, _P, _PP]),
    "vmx_rarr_from_raw": (C.c_int, [_P, _SZ, _P, _SZ, C.c_uint, _PP]),
    "vmx_rarr_prg_sha256": (C.c_int, [_P, _U8, _SZ, C.c_uint64, _SZ, C.c_uint, _PP]),
    "vmx_rarr_to_bytes": (C.c_int, [_P, _P]),
    "vmx_rarr_fill": (C.c_int, [_P, _SZ, _U8, _PP]),
    "vmx_rarr_free": (None, [_P]),
    "vmx_rarr_size": (_SZ, [_P]),
    "vmx_rarr_bitlen": (C.c_int, [_P, C.POINTER(C.c_uint)]),
    "vmx_radd": (C.c_int, [_P, _P, _PP]),
    "vmx_rneg": (C.c_int, [_P, _PP]),
    "vmx_rsub": (C.c_int, [_P, _P, _PP]),
    "vmx_rmul": (C.c_int, [_P, _P, _PP]),
    "vmx_rmuladd": (C.c_int, [_P, _U8, _P, _PP]),
    "vmx_rinner": (C.c_int, [_P, _P, _P]),
    "vmx_rsum": (C.c_int, [_P, _P]),
    "vmx_rprod": (C.c_int, [_P, _P]),
    "vmx_rprods": (C.c_int, [_P, _PP]),
    "vmx_rreclin": (C.c_int, [_P, _P, _PP, _P]),
    "vmx_rpermute": (C.c_int, [_P, _P, _PP]),
    "vmx_rshift_push": (C.c_int, [_P, _U8, _PP]),
    "vmx_rslice": (C.c_int, [_P, _SZ, _SZ, _PP]),
    "vmx_rget": (C.c_int, [_P, _SZ, _P]),
    "vmx_requals": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "vmx_ctx_launch_count": (C.c_uint64, [_P]),
    "vmx_ctx_modmul_count": (C.c_uint64, [_P]),
    "vmx_selftest_coop": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "vmx_selftest_sqr": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "vmx_debug_coop_mul": (C.c_int, [_P, _P, _PP]),
    "vmx_bench_modmul": (C.c_int, [_P, _SZ, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None
_lib_path = None


def library_path() -> str:
    return os.environ.get("VMX_LIBRARY_PATH", DEFAULT_LIB)


def load(path: str | None = None):
    """Load the engine.  Raises (never falls back) if the CUDA library is not built."""
    global _lib, _lib_path
    path = path or library_path()
    if _lib is not None and _lib_path == path:
        return _lib
    if not os.path.exists(path):
        raise VmxError(VMX_ECUDA, "engine library %s not built: run `python -c 'import __graft_entry__ as g; "
                       "g.build()'`; there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("VMX_TRACE"):  # development aid (bench.py --trace): record every ABI call with its times
        from ._trace import TracedLibrary
        lib = TracedLibrary(lib)
    _lib, _lib_path = lib, path
    return lib


def check(status: int) -> None:
    if status == VMX_OK:
        return
    msg = load().vmx_last_error().decode("utf-8", "replace")
    if status == VMX_EFORMAT:
        raise ArithmFormatError(status, msg)
    raise VmxError(status, msg)


def host_buffer(device: int, nbytes: int):
    """A pinned, pooled host buffer of `nbytes` bytes as a numpy uint8 array (include/vmx.h, "pinned host
    buffers"); it goes back to the pool when the last view of it is garbage-collected."""
    import weakref

    import numpy as np
    lib = load()
    p = C.c_void_p()
    check(lib.vmx_host_alloc(device, max(1, nbytes), C.byref(p)))
    raw = (C.c_uint8 * max(1, nbytes)).from_address(p.value)
    weakref.finalize(raw, lib.vmx_host_free, p.value)
    return np.frombuffer(raw, dtype=np.uint8, count=nbytes)
