"""`-m gpu`: sessions of other types and sessions after a pre-computation on the CUDA build (the bodies of
tests/test_engine_emul.py::test_mix_session_types and tests/test_vmnv_native.py::test_native_verifier_session_types;
the corrupted-directory variants are host logic and run in full on the CPU, here a handful)."""
import pytest

from tests import parity_bodies as pb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spec,n,mode,maxciph,width", [(2048, 30, "mixing", 45, 1), ("P-256", 8, "shuffling", 12, 2)])
def test_mix_session_types(engine_cuda, spec, n, mode, maxciph, width):
    """Pre-computation (permutation commitments + PoSC), keep lists, commitment-consistent shuffles and the
    decryption: proof directories byte-identical to the oracle's; the engine's vmnv and the oracle's agree
    (mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668)."""
    pb.mix_parity(engine_cuda, spec, n, width=width, mode=mode, maxciph=maxciph, light="min", gmp=isinstance(spec, int))


def test_native_vmnv_precomputed_session(engine_cuda):
    """libvmnv.so on a pre-computed session: PoSC + keep lists + CCPoS through vmx_* calls, same outcome as the mirror."""
    import __graft_entry__ as ge
    ge.build_vmnv()
    pb.native_vmnv_parity(engine_cuda, 3072, 6, mode="mixing", maxciph=9, minimal=True)
