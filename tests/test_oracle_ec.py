"""The oracle's P-256 group (oracle/ec.py): curve constants against their defining equations, group
laws, encodings round trip.  CPU only."""
import random

import pytest

from oracle import bytetree as bt
from oracle.crypto import SeededRandomSource
from oracle.ec import ECPoint, ECqPGroup, UNIT


def test_p256_constants_and_group_laws():
    G = ECqPGroup("P-256")
    assert G.p == 2 ** 256 - 2 ** 224 + 2 ** 192 + 2 ** 96 - 1 and G.a == G.p - 3
    assert G.on_curve(G.g) and G.op_exp(G.g, G.q).is_unit()
    # FIPS 186 / RFC test value: 2G
    two_g = G.op_mul(G.g, G.g)
    assert two_g.x == 0x7CF27B188D034F7E8A52380304B51AC3C08969E277F21B35A60B48FC47669978
    assert two_g.y == 0x07775510DB8ED040293D9AC69F7430DBBA7DADE63CE982299E04B79D227873D1
    rnd = random.Random(5)
    a, b, c = (rnd.randrange(G.q) for _ in range(3))
    A, B = G.op_exp(G.g, a), G.op_exp(G.g, b)
    assert G.op_mul(A, B) == G.op_exp(G.g, (a + b) % G.q)
    assert G.op_exp(A, c) == G.op_exp(G.g, a * c % G.q)
    assert G.op_mul(A, G.op_inv(A)) == UNIT and G.op_mul(A, UNIT) == A and G.op_exp(A, 0) == UNIT
    assert G.op_mul(A, A) == G.op_exp(A, 2)


def test_encodings_round_trip():
    G = ECqPGroup("P-256")
    rs = SeededRandomSource(bytes(range(32)))
    arr = G.random_array(7, rs, 100) + [UNIT]
    assert all(G.on_curve(P) and (P.is_unit() or P.y <= G.p - P.y) for P in arr)
    t = bt.from_bytes(G.leaf_array_tree(arr).to_bytes())
    assert G.parse_leaf_array(t, 8) == arr
    assert G.parse_leaf(bt.from_bytes(G.leaf_tree(arr[0]).to_bytes())) == arr[0]
    assert G.leaf_tree(UNIT).to_bytes().count(b"\xff" * 33) == 2
    bad = ECPoint(arr[0].x, (arr[0].y + 1) % G.p)
    try:
        G.parse_leaf(bt.from_bytes(G.leaf_tree(bad).to_bytes()))
        assert False
    except ValueError:
        pass


@pytest.mark.parametrize("curve", ["P-256", "secp256k1"])
def test_c_restatement_matches_python_oracle(curve):
    """oracle/cpu_ref_ec.c (the GMP-backed CPU baseline of the curve workloads) against oracle/ec.py: fixed-base,
    variable-base (per element and one exponent), simultaneous multiplication and point addition, with the
    unit element, equal and opposite operands and the exponents 0, 1, q - 1 among the inputs."""
    import random

    from oracle import accel, arithm as ar, ec as oec
    G = oec.ECqPGroup(curve)
    rnd = random.Random(11)
    n = 23
    es = [rnd.randrange(G.q) for _ in range(n)]
    es[0], es[1], es[2] = 0, 1, G.q - 1
    pts = [G.op_exp(G.g, rnd.randrange(1, G.q)) for _ in range(n)]
    pts[3] = G.one
    pts[5] = pts[4]
    pts[7] = G.op_inv(pts[6])
    rot = pts[1:] + pts[:1]
    want = (ar.g_exp(G, G.g, es), ar.g_exp(G, pts, es), ar.g_exp(G, pts, es[5]), ar.g_exp_prod(G, pts, es),
            ar.g_mul(G, pts, rot), ar.g_mul(G, pts, pts), ar.g_exp_prod(G, (pts, rot), es))
    undo = accel.install_ec(G, 3)
    try:
        got = (ar.g_exp(G, G.g, es), ar.g_exp(G, pts, es), ar.g_exp(G, pts, es[5]), ar.g_exp_prod(G, pts, es),
               ar.g_mul(G, pts, rot), ar.g_mul(G, pts, pts), ar.g_exp_prod(G, (pts, rot), es))
    finally:
        undo()
    assert got == want


def test_cpu_baseline_runs_on_the_curve():
    from oracle import cpu_baseline
    r = cpu_baseline.run(n_total=1000, sample=24, group="P-256")
    assert r["value"] > 0 and "cpu_ref_ec.c" in r["sample"]
