"""The oracle's P-256 group (oracle/ec.py): curve constants against their defining equations, group
laws, encodings round trip.  CPU only."""
import random

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
