"""`-m gpu`: sessions of every type and sessions after a pre-computation on the CUDA build (the bodies of
tests/test_engine_emul.py::test_mix_session_types and tests/test_vmnv_native.py::test_native_verifier_session_types)."""
import pytest

from tests import parity_bodies as pb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spec,n,mode,maxciph,width", [(2048, 6, "mixing", 9, 1), ("P-256", 12, "shuffling", 20, 2),
                                                       (3072, 5, "decryption", None, 1)])
def test_mix_session_types(engine_cuda, spec, n, mode, maxciph, width):
    """Proof directories byte-identical to the oracle's; the engine's vmnv and the oracle's agree on honest and
    corrupted directories (mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668, every type, pre-computation)."""
    pb.mix_parity(engine_cuda, spec, n, width=width, mode=mode, maxciph=maxciph, light=True)


@pytest.mark.parametrize("spec,n,mode,maxciph,width", [(2048, 8, "mixing", 12, 1), ("P-256", 9, "shuffling", 14, 2)])
def test_native_vmnv_session_types(engine_cuda, spec, n, mode, maxciph, width):
    """libvmnv.so on pre-computed sessions: PoSC + keep lists + CCPoS through vmx_* calls, same outcome as the mirror."""
    import __graft_entry__ as ge
    ge.build_vmnv()
    pb.native_vmnv_parity(engine_cuda, spec, n, width=width, mode=mode, maxciph=maxciph, thorough=False)
