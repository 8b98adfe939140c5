"""The oracle against every pinned vector available for the formats."""
import json
import os

from oracle import bytetree as bt
from oracle import gen_groups
from oracle.crypto import PRGHeuristic, RandomOracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_prg_kat():
    kat = json.load(open(os.path.join(GOLDEN, "prg_ro_kat.json")))
    prg = PRGHeuristic("sha256")
    prg.set_seed(bytes.fromhex(kat["seed_hex"]))
    assert prg.get_bytes(64).hex() == kat["prg_sha256_first_64_bytes_hex"]


def test_random_oracle_kat():
    kat = json.load(open(os.path.join(GOLDEN, "prg_ro_kat.json")))
    for bits, want in kat["ro_sha256"].items():
        assert RandomOracle("sha256", int(bits)).hash(bytes.fromhex(kat["seed_hex"])).hex() == want


def test_marshalled_modpgroup_fixture():
    """The one binary byte tree the reference ships (bench_config:43) parses under the restated
    rules: Marshalizer node(leaf(class name), node(p, q, g, leaf(be32 encoding))), leaves in
    minimal two's complement, p = 2q + 1, g of order q; and equals the explicit group of
    group_descriptions:32."""
    d = json.load(open(os.path.join(GOLDEN, "modpgroup_bench_config.json")))
    raw = bytes.fromhex(d["marshalled_hex"])
    t = bt.from_bytes(raw)
    assert t.to_bytes() == raw
    assert t.children[0].value == b"com.verificatum.arithm.ModPGroup"
    pl, ql, gl, enc = t.children[1].children
    p, q, g = (bt.bytes_to_int(x.value) for x in (pl, ql, gl))
    assert p.bit_length() == 15492 and p == 2 * q + 1 and pow(g, q, p) == 1 and g not in (0, 1)
    assert len(pl.value) == bt.int_byte_length(p) == 1937
    assert bt.int_to_bytes(p) == pl.value and bt.int_to_bytes(q, 1937) == ql.value
    assert enc.value == b"\x00\x00\x00\x01"
    assert int(d["explicit_p_hex"], 16) == p and int(d["explicit_g_hex"], 16) == g


def test_bytetree_roundtrip_and_errors():
    t = bt.node(bt.leaf(b"abc"), bt.node(bt.int_leaf(-5), bt.int32_leaf(7)), bt.node())
    raw = t.to_bytes()
    assert bt.from_bytes(raw) == t
    assert raw[:5] == b"\x00\x00\x00\x00\x03"
    for bad in (raw[:-1], raw + b"\x00", b"\x02" + raw[1:], b""):
        try:
            bt.from_bytes(bad)
            assert False
        except bt.EIOError:
            pass
    assert bt.int_to_bytes(255) == b"\x00\xff" and bt.int_to_bytes(127) == b"\x7f" and bt.int_to_bytes(-1) == b"\xff"


def test_rfc3526_constants(vmx):
    groups = __import__("importlib").import_module("verificatum-vmn_b200.groups")
    for bits, c in gen_groups.GROUPS.items():
        p = gen_groups.modp(bits, c)
        assert groups.rfc3526(bits) == (p, (p - 1) // 2, 2)
        assert p % 8 == 7 and pow(2, (p - 1) // 2, p) == 1
    p, q, g = groups.test512()
    assert gen_groups.is_probable_prime(p) and gen_groups.is_probable_prime(q) and pow(g, q, p) == 1


def test_async_digest_matches_and_releases_its_worker():
    """crypto.AsyncDigest (Fiat-Shamir hashing beside the GPU): same value as the synchronous oracle, and a digest
    that is abandoned or simply dropped (an exception on the way) does not leave a thread behind."""
    import gc
    import importlib
    import threading
    import time
    cr = importlib.import_module("verificatum-vmn_b200.crypto")
    ro = cr.RandomOracle(cr.HashfunctionHeuristic("SHA-256"), 256)
    n0 = threading.active_count()
    d = cr.AsyncDigest(ro.getDigest())
    for piece in (b"abc", b"d" * 100000, b"ef"):
        d.update(piece)
    assert d.digest() == ro.hash(b"abc" + b"d" * 100000 + b"ef")
    d = cr.AsyncDigest(ro.getDigest())
    d.update(b"abc")
    d.abandon()
    d = cr.AsyncDigest(ro.getDigest())
    d.update(b"abc")
    del d
    gc.collect()
    time.sleep(0.3)
    assert threading.active_count() == n0
