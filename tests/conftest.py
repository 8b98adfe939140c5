"""pytest plumbing: `gpu` marker, import of the hyphenated package, engine backends.

Two builds of the same C ABI are exercised:
  * the product, verificatum-vmn_b200/libvmx.so (CUDA, sm_100a)           -> tests marked `gpu`
  * tests/host_emul/libvmx_emul.so (g++ -DVMX_HOST_EMUL: every kernel body run sequentially on
    the CPU, built by the `emul_lib` fixture)                                -> CPU tests of the HOST logic
The emulation build is test infrastructure; the package never loads it on its own.
"""
import importlib
import os
import sys

import pytest

# published messages of the small test instances take the pooled page-locked buffer path (read-only
# memoryviews) that full-size ones take (eio.ByteTreeBasic.to_buffer); read when the package is imported
os.environ.setdefault("VMX_BUFFER_MIN", "256")
# ... and their multi-exponentiations the length-sorted chunk order of full-size ones (csrc/vmx.cu, seg_product)
os.environ.setdefault("VMX_MEXP_SORT_MIN", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def vmx():
    return importlib.import_module("verificatum-vmn_b200")


@pytest.fixture(scope="session")
def emul_lib():
    # tools/asan_emul.sh points this at an AddressSanitizer / UBSan build of the same sources
    if os.environ.get("VMX_EMUL_LIBRARY"):
        return os.environ["VMX_EMUL_LIBRARY"]
    import __graft_entry__ as ge
    return ge.build_host_emul()


@pytest.fixture()
def engine_emul(vmx, emul_lib, monkeypatch):
    """The package bound to the host-emulation build of the C ABI."""
    monkeypatch.setenv("VMX_LIBRARY_PATH", emul_lib)
    yield vmx
    import gc
    gc.collect()


@pytest.fixture()
def engine_cuda(vmx, monkeypatch):
    """The package bound to the CUDA build (fails loudly if it is missing or no GPU is present)."""
    monkeypatch.delenv("VMX_LIBRARY_PATH", raising=False)
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    yield vmx
    import gc
    gc.collect()


def pkg(name: str):
    return importlib.import_module("verificatum-vmn_b200." + name)
