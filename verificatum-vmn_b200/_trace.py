"""Host-side timeline of one run (development aid for bench.py --trace; off unless enabled).

Every call through the C ABI (`_native.load()` returns a recording proxy while tracing is on) and every hashing
call is recorded as (name, start, end, thread); `bench.py --trace FILE` dumps the list so that the gaps in which
the GPU waits for the host (or the host for the GPU) can be read off."""
from __future__ import annotations

import threading
import time

enabled = False
events = []


def start() -> None:
    global enabled
    events.clear()
    enabled = True


def stop():
    global enabled
    enabled = False
    return list(events)


class span:
    __slots__ = ("name", "t0", "extra")

    def __init__(self, name: str, extra=None):
        self.name, self.extra = name, extra

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if enabled:
            events.append((self.name, self.t0, time.perf_counter(), threading.get_ident(), self.extra))
        return False


class TracedLibrary:
    """Attribute-for-attribute proxy of the ctypes library that records the calls while tracing is enabled."""

    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)

    def __getattr__(self, name):
        fn = getattr(self._lib, name)

        def call(*args):
            if not enabled:
                return fn(*args)
            t0 = time.perf_counter()
            r = fn(*args)
            events.append((name, t0, time.perf_counter(), threading.get_ident(), None))
            return r

        object.__setattr__(self, name, call)
        return call
