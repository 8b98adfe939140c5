"""eio.HostBytes -- the immutable bytes-like value large published messages are held in (a pooled page-locked
buffer behind it on the GPU box) -- behaves like `bytes` wherever the package, the tests and a user of
mixnet.ShuffleProof treat a message as bytes.  Pure host logic: no engine needed."""
import copy
import dataclasses
import hashlib
import importlib
import io
import pickle

import numpy as np

eio = importlib.import_module("verificatum-vmn_b200.eio")


def _hb(data: bytes):
    return eio.HostBytes(np.frombuffer(bytearray(data), dtype=np.uint8))


def test_behaves_like_bytes():
    raw = bytes(range(256)) * 3
    h = _hb(raw)
    assert h == raw and raw == h and h == _hb(raw) and not (h == raw[:-1]) and h != raw + b"x"
    assert len(h) == len(raw) and bytes(h) == raw and bytearray(h) == bytearray(raw)
    assert h[5] == raw[5] and h[10:20] == raw[10:20] and h[:-2] == raw[:-2]
    m = memoryview(h)
    assert m.readonly and m.nbytes == len(raw) and m[3] == raw[3]
    assert hashlib.sha256(h).digest() == hashlib.sha256(raw).digest()
    f = io.BytesIO()
    f.write(h)
    assert f.getvalue() == raw
    assert np.frombuffer(h, dtype=np.uint8).tobytes() == raw


def test_copies_and_pickles_as_a_value():
    h = _hb(b"published message")
    assert copy.copy(h) is h and copy.deepcopy(h) is h
    assert pickle.loads(pickle.dumps(h)) == b"published message"

    @dataclasses.dataclass
    class Proof:
        reply: bytes

    p = Proof(h)
    assert dataclasses.asdict(p) == {"reply": b"published message"}
    assert dataclasses.astuple(p) == (b"published message",)
    assert dataclasses.replace(p, reply=bytes(h)[:-1]).reply == b"published messag"


def test_reader_parses_it_without_copying():
    tree = eio.ByteTreeContainer(eio.ByteTreeLeaf(b"abc"), eio.ByteTreeLeaf(b"defg")).to_bytes()
    r = eio.ByteTreeReader(_hb(tree))
    assert r.buf.readonly and r.getRemaining() == 2
    assert r.getNextChild().read() == b"abc" and r.getNextChild().read() == b"defg"


def test_to_buffer_uses_the_factory_above_the_threshold(monkeypatch):
    made = []

    def factory(n):
        made.append(n)
        return np.empty(n, dtype=np.uint8)

    monkeypatch.setattr(eio, "_buffer_factory", factory)
    monkeypatch.setattr(eio, "_BUFFER_MIN", 64)
    small = eio.ByteTreeLeaf(b"x" * 10)
    large = eio.ByteTreeContainer(*[eio.ByteTreeLeaf(bytes([i]) * 40) for i in range(5)])
    assert isinstance(small.to_buffer(), bytes) and not made
    out = large.to_buffer()
    assert isinstance(out, eio.HostBytes) and made == [large.total_bytes()] and out == large.to_bytes()
