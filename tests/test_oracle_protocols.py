"""Behavioural pins of the oracle's protocol restatement: what the reference's own tests assert
(hvzk/TestPoSCBasicTW.java:147-163: an honest transcript is accepted, a corrupted one is
rejected), on the reference test's own scale (ModPGroup(512), small N)."""
import pytest

from oracle import arithm as oar
from oracle import bytetree as bt
from oracle import protocols as opr
from oracle.crypto import SeededRandomSource
from tests.cases import OracleCase, seed


@pytest.fixture(scope="module")
def case():
    return OracleCase(512, 12)


@pytest.fixture(scope="module")
def proof(case):
    rs = SeededRandomSource(seed("prover"))
    wp, pr = opr.shuffle_and_prove(case.G, case.params, case.pk, case.w, case.h, rs)
    return wp, pr


def test_accepting_transcript(case, proof):
    assert opr.verify_shuffle(case.G, case.params, case.pk, case.w, case.h, proof[1])


def test_output_is_a_reencryption_of_a_permutation(case, proof):
    """decrypt(w') is a permutation of decrypt(w)."""
    G, x = case.G, case.x
    dec = lambda w: sorted(v * pow(pow(u, x, G.p), -1, G.p) % G.p for u, v in zip(*w))
    assert dec(proof[0]) == dec(case.w)
    assert proof[0] != case.w


@pytest.mark.parametrize("field", ["reply", "commitment", "permutationCommitment", "output"])
def test_rejecting_transcripts(case, proof, field):
    """Flip one bit of one published message (TestPoSCBasicTW.rejectingTranscript corrupts r)."""
    pr = dict(proof[1])
    raw = bytearray(pr[field])
    raw[len(raw) // 2] ^= 0x01
    pr[field] = bytes(raw)
    assert not opr.verify_shuffle(case.G, case.params, case.pk, case.w, case.h, pr)


def test_reply_field_by_field(case, proof):
    """Doubling any single reply value (r = r.add(r) in the reference test) is rejected."""
    G = case.G
    t = bt.from_bytes(proof[1]["reply"])
    for idx in range(6):
        kids = list(t.children)
        c = kids[idx]
        if c.is_leaf():
            v = bt.bytes_to_int(c.value)
            kids[idx] = bt.int_leaf((2 * v + 1) % G.q, G.ring_bytes)
        else:
            sub = list(c.children)
            v = bt.bytes_to_int(sub[0].value)
            sub[0] = bt.int_leaf((2 * v + 1) % G.q, G.ring_bytes)
            kids[idx] = bt.node(sub)
        pr = dict(proof[1])
        pr["reply"] = bt.node(kids).to_bytes()
        assert not opr.verify_shuffle(G, case.params, case.pk, case.w, case.h, pr), idx


def test_malformed_trees_are_rejected_not_raised(case, proof):
    for field in ("reply", "commitment", "permutationCommitment", "output"):
        for bad in (b"", b"\x01\x00\x00\x00\x00", proof[1][field][:-3]):
            pr = dict(proof[1])
            pr[field] = bad
            assert opr.verify_shuffle(case.G, case.params, case.pk, case.w, case.h, pr) is False


def test_ring_scans():
    G = oar.ModPGroup(23, 11, 4)
    b, e = [3, 5, 7, 9], [2, 4, 6, 8]
    x, d = oar.r_rec_lin(G, b, e)
    assert x == [3, (3 * 4 + 5) % 11, ((3 * 4 + 5) * 6 + 7) % 11, (((3 * 4 + 5) * 6 + 7) * 8 + 9) % 11] and d == x[-1]
    assert oar.r_prods(G, e) == [2, 8, 48 % 11, 384 % 11]
    assert oar.permute([10, 11, 12], [2, 0, 1]) == [11, 12, 10]
    assert oar.perm_inv([2, 0, 1]) == [1, 2, 0]


def test_cpu_baseline_of_the_mix_verification_runs():
    """oracle/cpu_baseline.run_verify_mix (the CPU arm of `bench.py --workload verify-mix`): a 3-party mix on the
    GMP back end is produced and accepted, on the small test group and on the curve."""
    from oracle import cpu_baseline
    for kw in (dict(bits=512), dict(group="P-256")):
        r = cpu_baseline.run_verify_mix(n_total=100, sample=12, **kw)
        assert r["value"] > 0 and "3-party mix" in r["sample"]
