"""CPU tests of the HOST side of the engine (handle management, table construction, Pippenger
plan, segmented products, scans, codecs, PRG streams, protocol mirror) through the C ABI of the
host-emulation build, against the oracle.  The same bodies run on the CUDA build in
tests/test_gpu_parity.py."""
import pytest

from tests import parity_bodies as pb


@pytest.mark.parametrize("n", [3, 37, 300])
def test_group_ops(engine_emul, n):
    pb.group_ops(engine_emul, 512, n)


def test_edge_cases(engine_emul):
    pb.edge_cases(engine_emul, 512)


@pytest.mark.parametrize("n", [1, 33, 1100])
def test_ring_ops(engine_emul, n):
    pb.ring_ops(engine_emul, 512, n)


def test_random_sources(engine_emul):
    pb.random_sources(engine_emul, 512, 41)


@pytest.mark.parametrize("n", [1, 2, 25])
def test_transcript_parity(engine_emul, n):
    pb.transcript_parity(engine_emul, 512, n)


def test_variable_base_kernels_in_several_launches(engine_emul, monkeypatch):
    """VMX_VAR_CHUNK bounds the elements per launch of k_exp_var / k_exp_var2 (production: the table scratch
    bound, reached above ~244k elements at 3072 bits): the i0 > 0 launches, and Pippenger at c = 12."""
    pb.with_env(monkeypatch, VMX_VAR_CHUNK=3, VMX_MEXP_WINDOW=8)
    pb.group_ops(engine_emul, 512, 17)
    pb.decryption_parity(engine_emul, 512, 7, 3, 2)
    pb.transcript_parity(engine_emul, 512, 7)


def test_production_kernels_body_on_emulation(engine_emul):
    """The body of the GPU production-shape test (GMP oracle) at a size the emulation build follows."""
    pb.production_kernels(engine_emul, 512, 150, fixed_windows=(5, 9), mexp_window=8, var_chunk=64)


def test_accept_reject_larger(engine_emul):
    pb.accept_reject_properties(engine_emul, 512, 700)


def test_group_ops_2048_small(engine_emul):
    pb.group_ops(engine_emul, 2048, 5)


def test_ring_ops_2048(engine_emul):
    """Long residues take the 4-per-thread scan (several levels of recursion at n = 300)."""
    pb.ring_ops(engine_emul, 2048, 300)


@pytest.mark.parametrize("n", [1, 19])
def test_posc_parity(engine_emul, n):
    pb.posc_parity(engine_emul, 512, n)


@pytest.mark.parametrize("n", [1, 19])
def test_ccpos_parity(engine_emul, n):
    pb.ccpos_parity(engine_emul, 512, n)


@pytest.mark.parametrize("n,k,t", [(1, 1, 1), (17, 3, 2), (9, 5, 3)])
def test_decryption_parity(engine_emul, n, k, t):
    pb.decryption_parity(engine_emul, 512, n, k, t)


@pytest.mark.parametrize("maxciph,n", [(12, 12), (15, 7), (9, 1)])
def test_committed_shuffle_parity(engine_emul, maxciph, n):
    pb.committed_shuffle_parity(engine_emul, 512, maxciph, n)


# ---- ECqPGroup (P-256): the same host logic and protocol mirror over the curve engine
@pytest.mark.parametrize("n", [1, 2, 40])
def test_ec_group_ops(engine_emul, n):
    pb.ec_group_ops(engine_emul, "P-256", n)


def test_ec_ring_ops(engine_emul):
    pb.ring_ops(engine_emul, "P-256", 33)


@pytest.mark.parametrize("n", [1, 9])
def test_ec_transcript_parity(engine_emul, n):
    pb.transcript_parity(engine_emul, "P-256", n)


def test_ec_posc_ccpos_parity(engine_emul):
    pb.posc_parity(engine_emul, "P-256", 7)
    pb.ccpos_parity(engine_emul, "P-256", 7)


def test_ec_decryption_parity(engine_emul):
    pb.decryption_parity(engine_emul, "P-256", 6, 3, 2)


def test_ec_committed_shuffle_parity(engine_emul):
    pb.committed_shuffle_parity(engine_emul, "P-256", 9, 5)


def test_ec_other_curve(engine_emul):
    """secp256k1: a = 0 (general doubling formula) and the generic Montgomery reduction."""
    pb.ec_group_ops(engine_emul, "secp256k1", 5)


@pytest.mark.parametrize("spec,n", [(512, 9), ("P-256", 6)])
def test_mix_and_vmnv_parity(engine_emul, spec, n, tmp_path):
    pb.mix_parity(engine_emul, spec, n, tmpdir=tmp_path)


@pytest.mark.parametrize("curve", ["P-256", "secp256k1"])
def test_ec_edge_cases(engine_emul, curve):
    pb.ec_edge_cases(engine_emul, curve)


@pytest.mark.parametrize("spec,width,n", [(512, 3, 6), ('P-256', 2, 4)])
def test_wide_ciphertexts_parity(engine_emul, spec, width, n):
    """BASELINE.json config 4: width-3 ciphertexts, shuffle + PoS and pre-computation + CCPoS."""
    pb.wide_shuffle_parity(engine_emul, spec, width, n)
    pb.wide_committed_shuffle_parity(engine_emul, spec, width, n + 3, n)


@pytest.mark.parametrize("spec,width,n", [(512, 3, 5), ("P-256", 2, 4)])
def test_mix_and_vmnv_parity_wide(engine_emul, spec, width, n, tmp_path):
    """A whole mix of width-omega ciphertexts and its verification (any width, elgamal/ProtocolElGamal.java:769-800)."""
    pb.mix_parity(engine_emul, spec, n, tmpdir=tmp_path, width=width)


@pytest.mark.parametrize("mode,maxciph,width,light", [("mixing", 6, 1, True), ("shuffling", None, 1, True),
                                                      ("shuffling", 5, 2, True), ("decryption", None, 2, False)])
def test_mix_session_types(engine_emul, mode, maxciph, width, light):
    """Sessions of type "shuffling" and "decryption", and sessions after a pre-computation (proofs of shuffles of
    commitments, keep lists, commitment-consistent proofs of shuffles): proof directories byte-identical to the
    oracle's, same verdicts on honest and corrupted directories and under every option of vmnv
    (mixnet/MixNetElGamalSession.java:161-358, mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668)."""
    pb.mix_parity(engine_emul, 512, 3, width=width, mode=mode, maxciph=maxciph, light=light)


def test_malformed_proof_files_are_verdicts_not_crashes(engine_emul):
    pb.malformed_proof_files(engine_emul, 512, 3)


@pytest.mark.parametrize("bits,n", [(2048, 12), (3072, 8)])
def test_dedicated_squaring(engine_emul, bits, n):
    """mont_sqr_tri (4 and 6 blocks of 16 words) on the emulated carry chains."""
    pb.squaring_selftest(engine_emul, bits, n, iters=2)


def test_one_context_two_threads(engine_emul):
    pb.concurrent_threads(engine_emul, 512, 40)


def test_recycled_blocks_and_table_eviction(engine_emul, monkeypatch):
    """The allocator paths of a BASELINE-sized run at test size: every device block of 4 KB and more goes through the
    context's recycling list (production: 32 MB), at most 64 KB parked (so blocks are also evicted), and the
    fixed-base table cache holds one small table at a time (so every change of base evicts and rebuilds)."""
    pb.with_env(monkeypatch, VMX_BIG_BLOCK_MIN=4096, VMX_BIG_CACHE_MAX=65536, VMX_TABLE_BUDGET=1)
    pb.group_ops(engine_emul, 512, 37)
    pb.transcript_parity(engine_emul, 512, 9)
    pb.decryption_parity(engine_emul, 512, 9, 3, 2)
