"""The native universal verifier (libvmnv.so over the C ABI, include/vmnv.h) against the Python mirror of the
reference's vmnv on the host-emulation build; the same body runs on the B200 in tests/test_gpu_parity.py."""
import ctypes
import os
import re

import pytest

from tests import parity_bodies as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vmnv_lib():
    if os.environ.get("VMNV_LIBRARY_PATH"):   # tools/asan_emul.sh: the sanitizer build of the same source
        return os.environ["VMNV_LIBRARY_PATH"]
    import __graft_entry__ as ge
    return ge.build_vmnv()


def test_library_exports_what_the_header_declares(vmnv_lib):
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vmnv.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(vmxv_[a-z0-9_]+)\s*\(", src)))
    assert syms == ["vmxv_bind", "vmxv_verify"]
    lib = ctypes.CDLL(vmnv_lib)
    for s in syms:
        assert hasattr(lib, s)


def test_native_verifier_matches_the_mirror(engine_emul, vmnv_lib):
    pb.native_vmnv_parity(engine_emul, 512, 3)


def test_native_verifier_wide_ciphertexts_other_thresholds(engine_emul, vmnv_lib):
    pb.native_vmnv_parity(engine_emul, 512, 3, k=4, threshold=3, width=2, thorough=False)


def test_native_verifier_curve_group(engine_emul, vmnv_lib):
    """ECqPGroup proofs: points as node(x, y), arrays as node(x leaves, y leaves), on-curve checks by the engine."""
    pb.native_vmnv_parity(engine_emul, "P-256", 4, thorough=False)
    pb.native_vmnv_parity(engine_emul, "P-256", 3, width=2, thorough=False)


@pytest.mark.parametrize("mode,maxciph,width,thorough", [("mixing", 6, 1, False), ("shuffling", 5, 2, False),
                                                         ("shuffling", None, 1, False), ("decryption", None, 1, True)])
def test_native_verifier_session_types(engine_emul, vmnv_lib, mode, maxciph, width, thorough):
    """Proofs of type "shuffling" / "decryption" and proofs after a pre-computation (PoSC, keep lists, CCPoS), the
    options -nodec / -noposc / -noccpos and the expected type: same outcome as the mirror."""
    pb.native_vmnv_parity(engine_emul, 512, 3, width=width, mode=mode, maxciph=maxciph, thorough=thorough)


@pytest.mark.parametrize("k,threshold,n,mode,maxciph", [(1, 1, 2, "mixing", 2), (2, 1, 3, "shuffling", 3),
                                                        (3, 3, 1, "decryption", None)])
def test_edge_sizes(engine_emul, vmnv_lib, k, threshold, n, mode, maxciph):
    """A single party, one ciphertext, a pre-computation for exactly the number of ciphertexts (keep lists of ones), a
    threshold equal to the number of parties: oracle, mirror and libvmnv agree."""
    pb.mix_parity(engine_emul, 512, n, k=k, threshold=threshold, mode=mode, maxciph=maxciph, light="min")
    pb.native_vmnv_parity(engine_emul, 512, n, k=k, threshold=threshold, mode=mode, maxciph=maxciph, minimal=True)


def test_native_verifier_differential_fuzz(engine_emul, vmnv_lib):
    """A seeded sample of the differential fuzzer (tools/fuzz_vmnv.py runs thousands of rounds under ASan)."""
    rounds = int(os.environ.get("VMNV_FUZZ_ROUNDS", "15"))
    tally = pb.native_vmnv_fuzz(engine_emul, 512, 3, rounds=rounds)
    assert sum(tally.values()) == rounds and tally.get("failstop", 0) > 0
    tally = pb.native_vmnv_fuzz(engine_emul, 512, 3, rounds=rounds, mode="mixing", maxciph=5, seed_label="fuzz/precomp")
    assert sum(tally.values()) == rounds and tally.get("failstop", 0) > 0
