"""ORACLE (test infrastructure).  The reported CPU baseline: the oracle's restatement of
re-encryption + PoSBasicTW prove + verify (oracle/protocols.py) with its array operations on
GMP (oracle/cpu_ref.c: fixed-base tables, mpz_powm, simultaneous exponentiation -- the
algorithms of the reference's gmpmee/vmgj natives) over all host cores, on a bounded sample of
the bench workload.  A stand-in for the Java/GMP path (no JVM in the image); kind = "port"."""
from __future__ import annotations

import hashlib
import importlib
import time

from . import accel
from . import arithm as ar
from . import protocols as pr
from .crypto import SeededRandomSource


def run(bits: int = 3072, n_total: int = 100000, sample: int = 0, steps: int = 1, warmup: int = 0):
    groups = importlib.import_module("verificatum-vmn_b200.groups")  # constants only
    p, q, g = groups.rfc3526(bits) if bits != 512 else groups.test512()
    G = ar.ModPGroup(p, q, g)
    cores = accel.cores()
    if sample <= 0:
        # measured here: ~46 ms per ciphertext and core at 3072 bits; aim at ~15 s per step
        per_ct = 0.046 * (bits / 3072.0) ** 2
        sample = max(8 * cores, min(n_total, int(15.0 * cores / per_ct)))
    undo = accel.install(G, cores)
    try:
        params = pr.Params(pgroup_string="ModPGroup(RFC3526-%d)" % bits)
        rs = SeededRandomSource(hashlib.sha256(b"cpu-baseline/setup").digest())
        x = ar.ring_random_element(G, rs, 100)
        pk = (g, pow(g, x, p))
        w = pr.demo_ciphertexts(G, pk, sample, rs)
        h = pr.independent_generators(G, "sha256", params.prefix(), "generators", sample, params.rbitlen)
        times = []
        for i in range(warmup + steps):
            prs = SeededRandomSource(hashlib.sha256(b"cpu-baseline/step%d" % i).digest())
            t0 = time.time()
            wp, proof = pr.shuffle_and_prove(G, params, pk, w, h, prs)
            ok = pr.verify_shuffle(G, params, pk, w, h, proof)
            dt = time.time() - t0
            if not ok:
                raise RuntimeError("cpu baseline: verifier rejected an honest proof")
            if i >= warmup:
                times.append(dt)
        t = sum(times) / len(times)
    finally:
        undo()
    return {"value": sample / t, "ms_per_step": t * 1e3, "cores": cores,
            "sample": "%d of %d ciphertexts (re-encrypt + prove + verify incl. Fiat-Shamir hashing and Legendre-symbol "
                      "membership checks), GMP 6 via oracle/cpu_ref.c, %d threads" % (sample, n_total, cores)}
