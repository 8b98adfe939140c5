"""Host-side hashing: PRG, random oracle, random sources.

Mirror of `com.verificatum.crypto.{PRGHeuristic, RandomOracle, HashfunctionHeuristic,
RandomSource}` (verificatum-vcr 3.1.0) as used by hvzk/ChallengerRO.java:96-116,
hvzk/PoSBasicTW.java:533-538 and distr/IndependentGeneratorsRO.java:110-130.  Fiat-Shamir
hashing is a single SHA-256 stream per challenge and stays on the host (SURVEY.md §8a row
a18); array-sized PRG expansions run on the device (`vmx_rarr_prg_sha256`).
"""
from __future__ import annotations

import hashlib
import os
import queue
import struct
import threading
import time

from . import _trace


class HashfunctionHeuristic:
    def __init__(self, name: str = "SHA-256"):
        self.name = name
        self._py = name.lower().replace("-", "")
        self.output_bits = hashlib.new(self._py).digest_size * 8

    def getDigest(self):
        return hashlib.new(self._py)

    def hash(self, *parts: bytes) -> bytes:
        h = self.getDigest()
        for p in parts:
            h.update(p)
        return h.digest()


class RandomSource:
    def getBytes(self, n: int) -> bytes:
        raise NotImplementedError


class RandomDevice(RandomSource):
    """/dev/urandom (com.verificatum.crypto.RandomDevice)."""

    def getBytes(self, n: int) -> bytes:
        return os.urandom(n)


class PRGHeuristic(RandomSource):
    """PRG(H): H(seed || be32(0)) || H(seed || be32(1)) || ..."""

    def __init__(self, hashfunction: HashfunctionHeuristic | None = None):
        self.hf = hashfunction or HashfunctionHeuristic("SHA-256")
        self.seed = None
        self.counter = 0
        self.buf = bytearray()

    def minNoSeedBytes(self) -> int:
        return self.hf.output_bits // 8

    def setSeed(self, seed: bytes) -> None:
        if len(seed) < self.minNoSeedBytes():
            raise ValueError("seed too short")
        self.seed = bytes(seed)
        self.counter = 0
        self.buf = bytearray()

    def getBytes(self, n: int) -> bytes:
        blocks = []
        have = len(self.buf)
        while have < n:
            blocks.append(self.hf.hash(self.seed, struct.pack(">I", self.counter)))
            self.counter += 1
            have += len(blocks[-1])
        if blocks:
            self.buf += b"".join(blocks)
        out = bytes(self.buf[:n])
        del self.buf[:n]
        return out


# process-wide accounting of the Fiat-Shamir hashing (bytes fed to random-oracle digests and the seconds spent
# in their SHA-256 updates): bench.py reports the hashed bytes per step and the SHA-256 rate next to the
# end-to-end figure, so that its Amdahl bound (one stream per challenge) can be audited
_hash_account = {"bytes": 0, "seconds": 0.0}
_hash_lock = threading.Lock()


def hashed_bytes() -> int:
    return _hash_account["bytes"]


def hashed_seconds() -> float:
    return _hash_account["seconds"]


class RandomOracleDigest:
    def __init__(self, hf: HashfunctionHeuristic, out_bits: int):
        self.hf = hf
        self.out_bits = out_bits
        self.h = hf.getDigest()
        self.h.update(struct.pack(">I", out_bits))
        self.nbytes = 4

    def update(self, data) -> None:
        n = len(data) if not hasattr(data, "nbytes") else data.nbytes
        if n >= 1 << 16:
            t0 = time.perf_counter()
            if _trace.enabled:
                with _trace.span("sha256.update", n):
                    self.h.update(data)
            else:
                self.h.update(data)
            dt = time.perf_counter() - t0
            with _hash_lock:
                _hash_account["bytes"] += n
                _hash_account["seconds"] += dt
        else:
            self.h.update(data)
        self.nbytes += n

    def digest(self) -> bytes:
        prg = PRGHeuristic(self.hf)
        prg.setSeed(self.h.digest())
        out = bytearray(prg.getBytes((self.out_bits + 7) // 8))
        extra = (8 - self.out_bits % 8) % 8
        if extra:
            out[0] &= 0xFF >> extra
        return bytes(out)


class RandomOracle:
    """RO(H, n_out)(d) = leading n_out bits of PRG_H(H(be32(n_out) || d))."""

    def __init__(self, hashfunction: HashfunctionHeuristic, out_bits: int):
        self.hf = hashfunction
        self.out_bits = out_bits

    def getDigest(self) -> RandomOracleDigest:
        return RandomOracleDigest(self.hf, self.out_bits)

    def hash(self, data: bytes) -> bytes:
        d = self.getDigest()
        d.update(data)
        return d.digest()


def _digest_worker(q, inner, state) -> None:
    """Worker of an AsyncDigest.  It holds the queue and the inner digest, NOT the AsyncDigest itself, so that a
    digest dropped on an exception path is collected and its __del__ releases the worker."""
    while True:
        item = q.get()
        if item is None:
            return
        if state["err"] is None:
            try:
                inner.update(item)
            except BaseException as e:  # surfaced by digest()
                state["err"] = e


class AsyncDigest:
    """A RandomOracleDigest fed by one worker thread.

    Fiat-Shamir hashing is one SHA-256 stream per challenge (hvzk/ChallengerRO.java:96-116) and cannot be spread
    over cores, but it can run BESIDE the GPU: `update` only queues the buffer (hashlib releases the GIL while it
    hashes), so the caller goes on queueing kernels for everything that does not depend on the challenge.  The
    buffers must not change until `digest()` has returned (array serialisations are immutable)."""

    def __init__(self, inner: RandomOracleDigest):
        self.inner = inner
        self._q: "queue.SimpleQueue" = queue.SimpleQueue()
        self._state = {"err": None}
        self._t = threading.Thread(target=_digest_worker, args=(self._q, inner, self._state), daemon=True)
        self._t.start()

    def update(self, data) -> None:
        self._q.put(data)

    def digest(self) -> bytes:
        self._q.put(None)
        with _trace.span("digest.wait"):
            self._t.join()
        if self._state["err"] is not None:
            raise self._state["err"]
        return self.inner.digest()

    def __del__(self):
        # a digest dropped without digest()/abandon() (an exception on the way) must not leave its worker waiting
        try:
            self._q.put(None)
        except Exception:
            pass

    def abandon(self) -> None:
        """Stop the worker without waiting for it (a speculative hash whose premise failed)."""
        self._state["err"] = self._state["err"] or RuntimeError("abandoned")
        self._q.put(None)

    @property
    def nbytes(self) -> int:
        return self.inner.nbytes
