"""ORACLE (test infrastructure).  The reported CPU baseline: the oracle's restatement of
re-encryption + PoSBasicTW prove + verify (oracle/protocols.py) with its array operations on
GMP (oracle/cpu_ref.c: fixed-base tables, mpz_powm, simultaneous exponentiation -- the
algorithms of the reference's gmpmee/vmgj natives; oracle/cpu_ref_ec.c: the same over a prime
curve, the vec/vecj natives) over all host cores, on a bounded sample of
the bench workload.  A stand-in for the Java/GMP path (no JVM in the image); kind = "port"."""
from __future__ import annotations

import hashlib
import importlib
import time

from . import accel
from . import arithm as ar
from . import protocols as pr
from .crypto import SeededRandomSource


def default_sample(bits: int, n_total: int, group: str = "modp", mix: bool = False) -> int:
    """Ciphertexts in the bounded sample one CPU step runs (about 15 s of work on this host's cores)."""
    cores = accel.cores()
    if group != "modp":
        # measured: ~13 ms per ciphertext and core on P-256 (a third of it the Python orchestration of the oracle
        # around the GMP calls), 16 ms for the verification of a mix
        per_ct = 0.016 if mix else 0.0125
    else:
        # measured: ~46 ms per ciphertext and core at 3072 bits (55 ms for the verification of a mix)
        per_ct = (0.055 if mix else 0.046) * (bits / 3072.0) ** 2
    return max(8 * cores, min(n_total, int(15.0 * cores / per_ct)))


def run(bits: int = 3072, n_total: int = 100000, sample: int = 0, steps: int = 1, warmup: int = 0,
        group: str = "modp"):
    cores = accel.cores()
    if group != "modp":   # a curve group (BASELINE.json config 5): cpu_ref_ec.c
        from . import ec
        G = ec.ECqPGroup(group)
        if sample <= 0:
            sample = default_sample(bits, n_total, group)
        undo = accel.install_ec(G, cores)
        label, lib_note, member_note = "ECqPGroup(%s)" % group, "oracle/cpu_ref_ec.c", "on-curve"
    else:
        groups = importlib.import_module("verificatum-vmn_b200.groups")  # constants only
        p, q, g = groups.rfc3526(bits) if bits != 512 else groups.test512()
        G = ar.ModPGroup(p, q, g)
        if sample <= 0:
            sample = default_sample(bits, n_total, group)
        undo = accel.install(G, cores)
        label, lib_note, member_note = "ModPGroup(RFC3526-%d)" % bits, "oracle/cpu_ref.c", "Legendre-symbol"
    try:
        params = pr.Params(pgroup_string=label)
        rs = SeededRandomSource(hashlib.sha256(b"cpu-baseline/setup").digest())
        x = ar.ring_random_element(G, rs, 100)
        pk = (G.g, ar.g_exp(G, G.g, x))
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
    return {"value": sample / t, "ms_per_step": t * 1e3, "cores": cores, "sample_n": sample,
            "sample": "%d of %d ciphertexts (re-encrypt + prove + verify incl. Fiat-Shamir hashing and %s "
                      "membership checks), GMP 6 via %s, %d threads" % (sample, n_total, member_note, lib_note, cores)}


def run_verify_mix(bits: int = 3072, n_total: int = 100000, sample: int = 0, steps: int = 1, warmup: int = 0,
                   group: str = "modp", k: int = 3, threshold: int = 2):
    """The same baseline for `bench.py --workload verify-mix` (BASELINE.json config 3): the oracle's vmnv
    (protocols.verify_mix: `threshold` verifyPoS + the verification of the decryption, from the bytes of a proof
    directory) over a k-party mix produced, untimed, by protocols.run_mix."""
    cores = accel.cores()
    if group != "modp":
        from . import ec
        G = ec.ECqPGroup(group)
        per_ct, label, lib_note = 0.016, "ECqPGroup(%s)" % group, "oracle/cpu_ref_ec.c"
        install = accel.install_ec
    else:
        groups = importlib.import_module("verificatum-vmn_b200.groups")  # constants only
        p, q, g = groups.rfc3526(bits) if bits != 512 else groups.test512()
        G = ar.ModPGroup(p, q, g)
        per_ct, label, lib_note = 0.055 * (bits / 3072.0) ** 2, "ModPGroup(RFC3526-%d)" % bits, "oracle/cpu_ref.c"
        install = accel.install
    if sample <= 0:
        sample = default_sample(bits, n_total, group, mix=True)
    undo = install(G, cores)
    try:
        params = pr.Params(pgroup_string=label)
        rs = SeededRandomSource(hashlib.sha256(b"cpu-baseline/mix").digest())
        probe = SeededRandomSource(hashlib.sha256(b"cpu-baseline/mix").digest())
        pk = (G.g, ar.g_exp(G, G.g, ar.ring_random_element(G, probe, params.rbitlen)))   # the key run_mix will deal
        w = pr.demo_ciphertexts(G, pk, sample, SeededRandomSource(hashlib.sha256(b"cpu-baseline/mix-input").digest()))
        d, _plain = pr.run_mix(G, params, k, threshold, w, rs)
        times = []
        for i in range(warmup + steps):
            t0 = time.time()
            rep = pr.verify_mix(G, params, k, threshold, d)
            dt = time.time() - t0
            if not rep["accepted"]:
                raise RuntimeError("cpu baseline: the verifier rejected an honest mix")
            if i >= warmup:
                times.append(dt)
        t = sum(times) / len(times)
    finally:
        undo()
    return {"value": sample / t, "ms_per_step": t * 1e3, "cores": cores, "sample_n": sample,
            "sample": "%d of %d ciphertexts (verification of a %d-party mix, threshold %d, from its proof directory "
                      "incl. Fiat-Shamir hashing and membership checks), GMP 6 via %s, %d threads"
                      % (sample, n_total, k, threshold, lib_note, cores)}


def run_committed_shuffle(bits: int = 2048, n_total: int = 100000, sample: int = 0, steps: int = 1, warmup: int = 0,
                          group: str = "modp", width: int = 3):
    """The same baseline for `bench.py --workload committed-shuffle` (BASELINE.json config 4's protocol): after an
    untimed pre-computation for `sample` ciphertexts (protocols.precomp / shrink), one step = re-encryption factors for
    fresh exponents, re-encryption + permutation, the commitment-consistent proof of a shuffle and its verification
    (protocols.committed_shuffle / ccpos_verify), ModPGroup only."""
    if group != "modp":
        raise ValueError("the CPU arm of the committed shuffle is written for ModPGroup")
    cores = accel.cores()
    groups = importlib.import_module("verificatum-vmn_b200.groups")  # constants only
    p, q, g = groups.rfc3526(bits) if bits != 512 else groups.test512()
    G = ar.ModPGroup(p, q, g)
    if sample <= 0:
        sample = max(8 * cores, min(n_total, int(15.0 * cores / (0.02 * width * (bits / 3072.0) ** 2))))
    undo = accel.install(G, cores)
    try:
        params = pr.Params(pgroup_string="ModPGroup(RFC3526-%d)" % bits)
        rs = SeededRandomSource(hashlib.sha256(b"cpu-baseline/committed").digest())
        pk = pr.wide_key((G.g, ar.g_exp(G, G.g, ar.ring_random_element(G, rs, 100))), width)

        def exponents(src):
            cols = tuple(ar.ring_random_array(G, sample, src, params.rbitlen) for _ in range(width))
            return cols if width > 1 else cols[0]
        w = ar.g_exp(G, pk, exponents(rs))
        h = pr.independent_generators(G, "sha256", params.prefix(), "generators", sample, params.rbitlen)
        state, _pub = pr.precomp(G, params, pk, h, rs)
        pr.shrink(G, state, sample)
        times = []
        for i in range(warmup + steps):
            prs = SeededRandomSource(hashlib.sha256(b"cpu-baseline/committed/step%d" % i).digest())
            t0 = time.time()
            state["s"] = exponents(prs)
            state["factors"] = ar.g_exp(G, pk, state["s"])
            wp, proof = pr.committed_shuffle(G, params, pk, state, w, prs)
            ok = pr.ccpos_verify(G, params, G.g, state["h"], state["u"], pk, w, wp, proof["commitment"], proof["reply"])
            dt = time.time() - t0
            if not ok:
                raise RuntimeError("cpu baseline: the verifier rejected an honest commitment-consistent proof")
            if i >= warmup:
                times.append(dt)
        t = sum(times) / len(times)
    finally:
        undo()
    return {"value": sample / t, "ms_per_step": t * 1e3, "cores": cores, "sample_n": sample,
            "sample": "%d of %d ciphertexts of width %d (re-encryption factors, re-encryption, CCPoS prove + verify incl. "
                      "Fiat-Shamir hashing), GMP 6 via oracle/cpu_ref.c, %d threads" % (sample, n_total, width, cores)}
