"""The drop-in boundary: the CUDA library loads without a GPU and exports every symbol that
include/vmx.h declares; the ctypes binding declares exactly that set; the product never
falls back to the CPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vmx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vmx_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def cuda_lib():
    import __graft_entry__ as ge
    return ge.build_engine()


def test_library_exports_every_declared_symbol(cuda_lib):
    lib = ctypes.CDLL(cuda_lib)
    syms = header_symbols()
    assert len(syms) > 50
    for s in syms:
        assert hasattr(lib, s), "include/vmx.h declares %s but libvmx.so does not export it" % s


def test_binding_matches_header(vmx):
    assert sorted(vmx._native.SIGNATURES) == header_symbols()


def test_no_cpu_fallback_without_device(vmx, cuda_lib, monkeypatch):
    """Without a CUDA device context creation fails with VMX_ECUDA -- it never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    monkeypatch.delenv("VMX_LIBRARY_PATH", raising=False)
    from tests.cases import group_params
    p, q, g = group_params(512)
    with pytest.raises(vmx._native.VmxError) as ei:
        vmx.arithm.ModPGroup(p, q, g)
    assert ei.value.status == vmx._native.VMX_ECUDA


def test_missing_library_fails_loudly(vmx, monkeypatch, tmp_path):
    monkeypatch.setenv("VMX_LIBRARY_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(vmx._native.VmxError):
        vmx._native.load()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "verificatum-vmn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "import_module(\"oracle" not in text and "libgmp" not in text, f
