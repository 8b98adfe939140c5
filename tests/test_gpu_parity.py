"""GPU parity tests: the CUDA engine behind the C ABI against the oracle, bit for bit, on the
groups of BASELINE.json (2048- and 3072-bit safe primes) and on the reference unit test's own
scale (ModPGroup(512)); plus size-independent properties at sizes the oracle cannot follow."""
import ctypes as C

import pytest

from tests import parity_bodies as pb
from tests.cases import group_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits,n", [(512, 300), (512, 9000), (2048, 40), (3072, 40)])
def test_group_ops(engine_cuda, bits, n):
    pb.group_ops(engine_cuda, bits, n)


@pytest.mark.parametrize("bits", [512, 2048, 3072])
def test_edge_cases(engine_cuda, bits):
    pb.edge_cases(engine_cuda, bits)


@pytest.mark.parametrize("bits,n", [(512, 1100), (3072, 70)])
def test_ring_ops(engine_cuda, bits, n):
    pb.ring_ops(engine_cuda, bits, n)


@pytest.mark.parametrize("bits", [512, 3072])
def test_random_sources(engine_cuda, bits):
    pb.random_sources(engine_cuda, bits, 41)


@pytest.mark.parametrize("bits,n", [(512, 1), (512, 100), (2048, 12), (3072, 12)])
def test_transcript_parity(engine_cuda, bits, n):
    """512/100 is the reference's hvzk/TestPoSCBasicTW.java scale; 2048 and 3072 are BASELINE.json's groups."""
    pb.transcript_parity(engine_cuda, bits, n)


@pytest.mark.parametrize("bits,n", [(512, 100), (3072, 10)])
def test_posc_parity(engine_cuda, bits, n):
    """512/100 is exactly hvzk/TestPoSCBasicTW.java's instance size."""
    pb.posc_parity(engine_cuda, bits, n)


@pytest.mark.parametrize("bits,n", [(512, 100), (2048, 10)])
def test_ccpos_parity(engine_cuda, bits, n):
    pb.ccpos_parity(engine_cuda, bits, n)


@pytest.mark.parametrize("bits,n,k,t", [(512, 60, 3, 2), (3072, 9, 3, 2), (2048, 8, 5, 3)])
def test_decryption_parity(engine_cuda, bits, n, k, t):
    pb.decryption_parity(engine_cuda, bits, n, k, t)


@pytest.mark.parametrize("bits,maxciph,n", [(512, 150, 100), (3072, 15, 10)])
def test_committed_shuffle_parity(engine_cuda, bits, maxciph, n):
    """precomp(15) + committedShuffle of 10 is the reference's mixnet/DemoShufflerElGamal.java:163-265 scale."""
    pb.committed_shuffle_parity(engine_cuda, bits, maxciph, n)


@pytest.mark.parametrize("bits", [512, 2048, 3072])
def test_cooperative_multiplier_matches_thread_per_element(engine_cuda, bits):
    vmx = engine_cuda
    A = vmx.arithm
    p, q, g = group_params(bits)
    G = A.ModPGroup(p, q, g)
    rs = vmx.crypto.PRGHeuristic()
    rs.setSeed(bytes(range(32)))
    X1 = G.randomElementArray(6000, rs, 100)
    X2 = G.randomElementArray(6000, rs, 100)
    eq = C.c_int()
    lib = vmx._native.load()
    vmx._native.check(lib.vmx_selftest_coop(X1.h, X2.h, C.byref(eq)))
    assert eq.value == 1
    edge = G.toElementArray([A.PGroupElement(G, v) for v in (1, p - 1, p - 2, g)])
    vmx._native.check(lib.vmx_selftest_coop(edge.h, edge.h, C.byref(eq)))
    assert eq.value == 1


@pytest.mark.parametrize("bits,n", [(3072, 10240), (2048, 12288)])
def test_production_kernels_against_gmp_oracle(engine_cuda, bits, n):
    """n above the cooperative-kernel bound (8192): k_exp_var / k_exp_var2 / k_exp_fixed at w = 16, 17 / Pippenger
    c = 12 / k_inv_up,down at 96 and 64 limbs, bit for bit against GMP."""
    pb.production_kernels(engine_cuda, bits, n)


def test_production_kernels_multi_launch(engine_cuda):
    """The variable-base kernels cut an array in several launches once their table scratch is bounded (n above
    ~244k at 3072 bits); forced here at n = 9000 with 4096 elements per launch (i0 > 0 offsets)."""
    pb.production_kernels(engine_cuda, 3072, 9000, fixed_windows=(16,), var_chunk=4096)


@pytest.mark.parametrize("bits,n", [(2048, 12), (3072, 11)])
def test_transcripts_through_thread_per_element_kernels(engine_cuda, monkeypatch, bits, n):
    """The oracle-sized protocol transcripts with the kernel selection of a BASELINE-sized run: no
    warp-cooperative shortcut for small arrays (VMX_COOP_MAX=0), several launches per array (VMX_VAR_CHUNK),
    Pippenger at c = 12: PoS prove/verify, decryption factors and their proof, byte for byte."""
    pb.with_env(monkeypatch, VMX_COOP_MAX=0, VMX_VAR_CHUNK=5, VMX_MEXP_WINDOW=12)
    pb.transcript_parity(engine_cuda, bits, n)
    pb.decryption_parity(engine_cuda, bits, n - 2, 3, 2)
    pb.group_ops(engine_cuda, bits, n + 9)


@pytest.mark.parametrize("bits,n", [(2048, 3000), (3072, 3000)])
def test_dedicated_squaring(engine_cuda, bits, n):
    pb.squaring_selftest(engine_cuda, bits, n, iters=5)


def test_accept_reject_at_scale_3072(engine_cuda):
    """N = 50,000 at 3072 bits (thread-per-element kernels, several waves, Pippenger c = 12)."""
    pb.accept_reject_properties(engine_cuda, 3072, 50000)


def test_accept_reject_at_scale_2048(engine_cuda):
    pb.accept_reject_properties(engine_cuda, 2048, 60000)


@pytest.mark.parametrize("bits,n", [(512, 3000), (3072, 300)])
def test_one_context_two_threads(engine_cuda, bits, n):
    """Export thread + compute thread on one context (hvzk/CCPoSW.java:116-122, mixnet/ShufflerElGamalSession.java:847-856)."""
    pb.concurrent_threads(engine_cuda, bits, n)


def test_launches_are_counted(engine_cuda):
    vmx = engine_cuda
    A = vmx.arithm
    p, q, g = group_params(512)
    G = A.ModPGroup(p, q, g)
    before = G.launch_count()
    rs = vmx.crypto.PRGHeuristic()
    rs.setSeed(bytes(range(32)))
    e = G.getPRing().randomElementArray(100, rs, 100)
    G.getg().exp(e).free()
    assert G.launch_count() > before and G.modmul_count() > 0


# ---- ECqPGroup (P-256, BASELINE.json config 5)
@pytest.mark.parametrize("curve,n", [("P-256", 1), ("P-256", 40), ("P-256", 3000), ("secp256k1", 33)])
def test_ec_group_ops(engine_cuda, curve, n):
    """n = 3000 crosses the block size of the batched inversion (512 points) and its recursion."""
    pb.ec_group_ops(engine_cuda, curve, n)


def test_ec_ring_ops(engine_cuda):
    pb.ring_ops(engine_cuda, "P-256", 1100)


@pytest.mark.parametrize("n", [1, 100])
def test_ec_transcript_parity(engine_cuda, n):
    pb.transcript_parity(engine_cuda, "P-256", n)


def test_ec_posc_ccpos_parity(engine_cuda):
    pb.posc_parity(engine_cuda, "P-256", 60)
    pb.ccpos_parity(engine_cuda, "P-256", 60)


def test_ec_decryption_parity(engine_cuda):
    pb.decryption_parity(engine_cuda, "P-256", 40, 3, 2)


def test_ec_committed_shuffle_parity(engine_cuda):
    pb.committed_shuffle_parity(engine_cuda, "P-256", 60, 40)


def test_ec_accept_reject_at_scale(engine_cuda):
    """N = 300,000 points: tables of 16-bit windows, Pippenger c = 16, several levels of the batched inversion."""
    pb.accept_reject_properties(engine_cuda, "P-256", 300000)


@pytest.mark.parametrize("spec,n", [(3072, 5), (512, 300), ("P-256", 80)])
def test_mix_and_vmnv_parity(engine_cuda, spec, n, tmp_path):
    """A 3-party mix with threshold 2 and its vmnv-style verification (BASELINE.json config 3 at oracle size)."""
    # (the oracle's array operations on its GMP back end for the ModP cases: the 512-bit mix of 300 took 36 s of
    # Python pow() with the GPU idle)
    pb.mix_parity(engine_cuda, spec, n, tmpdir=tmp_path, gmp=isinstance(spec, int))


@pytest.mark.parametrize("bits,n,width", [(3072, 12, 1), (2048, 40, 2)])
def test_native_vmnv(engine_cuda, bits, n, width):
    """libvmnv.so (the native universal verifier over the C ABI) on the CUDA build: same verdicts as the Python
    mirror on honest and corrupted proof directories."""
    import __graft_entry__ as ge
    ge.build_vmnv()
    pb.native_vmnv_parity(engine_cuda, bits, n, width=width, thorough=False)


@pytest.mark.parametrize("curve", ["P-256", "secp256k1"])
def test_ec_edge_cases(engine_cuda, curve):
    pb.ec_edge_cases(engine_cuda, curve)


@pytest.mark.parametrize("spec,width,n", [(2048, 3, 16), ("P-256", 3, 120)])
def test_wide_ciphertexts_parity(engine_cuda, spec, width, n):
    """BASELINE.json config 4: width-3 ciphertexts, shuffle + PoS and pre-computation + CCPoS."""
    pb.wide_shuffle_parity(engine_cuda, spec, width, n)
    pb.wide_committed_shuffle_parity(engine_cuda, spec, width, n + 3, n)
