#!/usr/bin/env python
"""bench.py -- ciphertexts/s of re-encryption + proof of shuffle (prove and verify) on B200.

One step = one pass of the hot path over one batch of N synthetic ciphertexts (3072-bit
ModPGroup, width 1): El Gamal re-encryption + permutation, PoSBasicTW commit/reply, and
PoSBasicTW verification of that proof.

  value  device-timed (CUDA events on the engine's stream, max over ranks): inputs, proof arrays
         and challenges resident in HBM; protocol through PoSBasicTW with given seed/challenge
         (what hvzk/TestPoSCBasicTW.java drives), no byte-tree traffic.
  e2e    the same work through the public API a mix-server calls (mixnet.ShufflerSession:
         shuffle -> ShuffleProof bytes -> verify) with HOST byte buffers: byte-tree
         encode/decode, H2D/D2H copies and the Fiat-Shamir SHA-256 hashing are inside the timed
         region.
  roofline  IMAD-pipe modmul roofline of the dominant kernel (fixed-base exponentiation).
  cpu_baseline  the oracle's GMP-backed C restatement on the host cores (bounded sample).

Other workloads (`--workload`): verify-mix = vmnv over the proof directory of a 3-party mix (BASELINE.json config 3);
committed-shuffle = pre-computation, then re-encryption + commitment-consistent proof of a shuffle, prove + verify
(config 4's protocol).  The default run reports configs 3, 4, 5 as `other_configs`.

`--impl reference` times that CPU restatement alone (the reference is Java + GMP natives; no JVM
exists in this image) on the same config.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# measured on B200 by tools/ubench.cu (profiles/r01_ubench_imad.txt): IMAD.WIDE.U32 issues at
# 0.99 warp-instr/clk/SM -> 32 MAC/clk/SM * 148 SM * 1.965 GHz
IMAD_PEAK_MAC_PER_S = 9.26e12


def macs_per_modmul(bits: int) -> int:
    n = bits // 32
    return 2 * n * n + n


def nominal_modmuls_per_ciphertext(bits: int, n: int, n_e=256, n_v=256, n_r=100, width=1):
    """SURVEY.md §8d textbook costs: FIX(L)=ceil(L/8), VAR(L)=L+ceil(L/6), MEXP(L)=min_c ceil(L/c)(1+2^(c+1)/N)."""
    Lq = bits - 1
    fix = lambda L: -(-L // 8)
    var = lambda L: L + -(-L // 6)
    mexp = lambda L: min(-(-L // c) * (1 + 2 ** (c + 1) / n) for c in range(1, 24))
    eps = n_e + n_v + n_r
    k = 1 + 2 * width
    reenc = 2 * width * (fix(Lq) + 1)
    prove = 5 * fix(Lq) + k * mexp(eps) + 4
    verify = k * mexp(n_e) + k * mexp(eps + 1) + var(n_v) + var(eps + 1) + fix(Lq) + 4
    return {"reencrypt": reenc, "prove": prove, "verify": verify, "total": reenc + prove + verify}


def nominal_fieldmuls_per_ciphertext_ec(n: int, L=256, width=1):
    """The same textbook operation counts on a 256-bit curve, in field multiplications (136 word MACs each,
    SURVEY.md §8d): mixed addition 11, doubling 8, full addition 16; every exponent is < q (256 bits)."""
    fix = lambda: -(-L // 8) * 11
    var = lambda: L * 8 + -(-L // 6) * 16
    mexp = lambda: min(-(-L // c) * (11 + 16 * 2 ** (c + 1) / n) for c in range(1, 24))
    k = 1 + 2 * width
    reenc = 2 * width * (fix() + 11)
    prove = 5 * fix() + k * mexp() + 4 * 11
    verify = 2 * k * mexp() + 2 * var() + fix() + 4 * 11
    return {"reencrypt": reenc, "prove": prove, "verify": verify, "total": reenc + prove + verify}


class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.sm_max = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def run_reference(args):
    """The reference arm: the oracle's CPU restatement (GMP) on the host cores."""
    from oracle import cpu_baseline
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "committed-shuffle":
        res = cpu_baseline.run_committed_shuffle(bits=args.bits, n_total=args.n, sample=cpu_sample_size(args),
                                                 steps=args.steps, warmup=min(args.warmup, 1), group=args.group,
                                                 width=args.width)
        line = {"metric": "ciphertexts/s: re-encrypt + CCPoS prove+verify after pre-computation, %s, width %d"
                          % (group_label(args), args.width), "impl": "reference",
                "value": res["value"], "unit": "ciphertexts/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32 limbs (exact integer)", "data": "synthetic",
                "config": {"workload": "%s, width %d, N=%d ciphertexts: re-encryption + commitment-consistent proof of "
                                       "a shuffle, prove + verify, after a pre-computation for N (BASELINE.json config 4's "
                                       "protocol)" % (group_label(args), args.width, args.n), "n_total": args.n,
                           "width": args.width, "cpu_sample": res["sample_n"]},
                "cpu_baseline": {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"], "kind": "port",
                                 "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "ciphertexts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    if args.workload == "verify-mix":
        res = cpu_baseline.run_verify_mix(bits=args.bits, n_total=args.n, sample=cpu_sample_size(args), steps=args.steps,
                                          warmup=min(args.warmup, 1), group=args.group)
    else:
        res = cpu_baseline.run(bits=args.bits, n_total=args.n, sample=cpu_sample_size(args), steps=args.steps,
                               warmup=min(args.warmup, 1), group=args.group)
    mixw = args.workload == "verify-mix"
    cfg = mix_config_dict(args, args.gpus) if mixw else config_dict(args, args.gpus)
    assert cfg["cpu_sample"] == res["sample_n"]
    line = {"metric": mix_metric_name(args) if mixw else metric_name(args), "impl": "reference",
            "value": res["value"], "unit": "ciphertexts/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 limbs (exact integer)", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"], "kind": "port",
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "ciphertexts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_sample_size(args) -> int:
    """Ciphertexts one step of the CPU arm runs: a bounded sample of the workload, about 15 s of work on this
    host's cores (measured per ciphertext and core: ~46 ms at 3072 bits, 55 ms for the verification of a mix;
    ~13 / 16 ms on P-256).  Both arms name it in `config`."""
    if args.cpu_sample:
        return args.cpu_sample
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    mix = args.workload == "verify-mix"
    if args.workload == "committed-shuffle":
        per_ct = 0.02 * args.width * (args.bits / 3072.0) ** 2
    elif is_curve(args):
        per_ct = 0.016 if mix else 0.0125
    else:
        per_ct = (0.055 if mix else 0.046) * (args.bits / 3072.0) ** 2
    return max(8 * cores, min(args.n, int(15.0 * cores / per_ct)))


def is_curve(args) -> bool:
    return args.group != "modp"


def group_label(args) -> str:
    return "ECqPGroup %s" % args.group if is_curve(args) else "ModPGroup %d-bit (RFC 3526)" % args.bits


def metric_name(args) -> str:
    return "ciphertexts/s: re-encrypt+PoS prove+verify, " + \
        ("%s ECqPGroup" % args.group if is_curve(args) else "%d-bit ModPGroup" % args.bits)


def mix_metric_name(args) -> str:
    return "ciphertexts/s: vmnv-style verification of a 3-party mix (2 x PoS + decryption proofs), " + \
        ("%s ECqPGroup" % args.group if is_curve(args) else "%d-bit ModPGroup" % args.bits)


def mix_config_dict(args, world=1):
    return dict(config_dict(args, world), workload="%s, width 1, N=%d ciphertexts in total (one list, sharded over "
                "%d GPU%s): verification of a 3-party mix with threshold 2 from its proof directory in host memory"
                % (group_label(args), args.n, world, "" if world == 1 else "s"), k=3, threshold=2)


def config_dict(args, world=1):
    elem = 64 if is_curve(args) else args.bits // 8
    return {"workload": "%s, width %d, N=%d ciphertexts in total (one list, sharded over %d GPU%s): re-encrypt + "
                        "PoSBasicTW prove + verify" % (group_label(args), args.width, args.n, world,
                                                       "" if world == 1 else "s"),
            "group": args.group, "bits": 256 if is_curve(args) else args.bits, "width": args.width,
            "n_total": args.n, "n_per_gpu": -(-args.n // world),
            "ebitlen": 256, "vbitlen": 256, "rbitlen": 100,
            # the CPU arm (cpu_baseline, --impl reference) runs a bounded SAMPLE of the N ciphertexts per step and
            # reports sample / time: this many, on all host cores
            "cpu_sample": cpu_sample_size(args),
            "l2": "working set (N x %d B per array, >10 arrays) exceeds the 126 MB L2" % elem}


class Env:
    """What every measurement of this process shares: torch, the process group, this rank."""

    def __init__(self, torch, dist, world, rank, local_rank):
        self.torch, self.dist, self.world, self.rank, self.local_rank = torch, dist, world, rank, local_rank

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def make_group(args, env):
    vmx = importlib.import_module("verificatum-vmn_b200")
    groups = importlib.import_module("verificatum-vmn_b200.groups")
    A = vmx.arithm
    if env.world > 1:
        # ONE list of n ciphertexts, sharded in contiguous index ranges over the GPUs: one shuffle, one proof;
        # expProd partial products and permuted rows travel over NCCL
        par = importlib.import_module("verificatum-vmn_b200.parallel")
        if is_curve(args):
            return par.make_curve_group(args.group, env.local_rank)
        return par.make_group(*groups.rfc3526(args.bits), env.local_rank)
    if is_curve(args):
        return A.ECqPGroup(args.group, device=env.local_rank)
    return A.ModPGroup(*groups.rfc3526(args.bits), device=env.local_rank)


def make_prg(label: str):
    # the same stream on every rank: each rank expands its own slice of it on the device
    crypto = importlib.import_module("verificatum-vmn_b200.crypto")
    r = crypto.PRGHeuristic()
    r.setSeed(crypto.HashfunctionHeuristic("SHA-256").hash(("vmx-bench/%s" % label).encode()))
    return r


def run_verify_mix(args, env):
    """BASELINE.json config 3: what `vmnv` does with the proof directory of a 3-party mix (threshold 2): two
    verifyPoS and the verification of the decryption (3 arrays of decryption factors, batched proof, plaintexts)
    -- mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668.  The proof directory is produced by the engine's
    own mix (untimed) and held in HOST memory; one step = one full verification from those bytes (import with
    membership checks, Fiat-Shamir hashing, all array work), so `value` and `e2e` are the same measurement here."""
    torch, world, rank = env.torch, env.world, env.rank
    vm = importlib.import_module("verificatum-vmn_b200.vmnv")
    mixnet = importlib.import_module("verificatum-vmn_b200.mixnet")
    G = make_group(args, env)
    stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx), device=torch.device("cuda", env.local_rank))
    n = args.n
    n_local = n * (rank + 1) // world - n * rank // world
    params = mixnet.SessionParams(pGroupString="%s" % group_label(args))
    M = vm.MixNetElGamal(G, params, 3, 2, make_prg("mix/dealer"))
    w = mixnet.demoCiphertexts(M.fullPublicKey, n, make_prg("mix/input"))
    t0 = time.time()
    M.run(w).free()
    G.sync()
    prove_s = time.time() - t0
    nizkp = M.nizkp
    V = vm.MixNetElGamalVerifyFiatShamirSession(G, params, 3, 2)
    G.membership_check = True

    for _ in range(max(1, min(args.warmup, 2))):
        if not V.verify(nizkp)["accepted"]:
            raise SystemExit("bench: the verifier rejected an honest mix")
    if args.trace:   # every rank runs the step (it is full of collectives); rank 0 keeps its timeline
        tr = importlib.import_module("verificatum-vmn_b200._trace")
        tr.start()
        with tr.span("e2e.step"):
            V.verify(nizkp)
        events = tr.stop()
        if rank == 0:
            with open(args.trace, "w") as f:
                json.dump(events, f)
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    env.barrier()
    launches0, modmuls0 = G.launch_count(), G.modmul_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.time()
    for _ in range(args.steps):
        V.verify(nizkp)
    e1.record(stream)
    e1.synchronize()
    env.barrier()
    wall = env.max_over_ranks(time.time() - t0)
    sampler.stop_flag.set()
    sampler.join()
    ms_per_step = env.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    line = None
    if rank == 0:
        macs = 136 if is_curve(args) else macs_per_modmul(args.bits)
        nbytes = sum(len(v) for v in nizkp.values())
        modmuls = G.modmul_count() - modmuls0
        line = {"metric": mix_metric_name(args),
                "value": n / (ms_per_step * 1e-3), "unit": "ciphertexts/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32 limbs (exact integer)", "data": "synthetic",
                "config": mix_config_dict(args, world),
                "clocks": sampler.summary(), "gpu_launches": int(G.launch_count() - launches0),
                "e2e": {"value": n / (wall / args.steps), "unit": "ciphertexts/s", "h2d_bytes_per_step": nbytes,
                        "d2h_bytes_per_step": 0, "ms_per_step": wall / args.steps * 1e3,
                        "includes": "byte-tree decode, H2D, membership checks, Fiat-Shamir SHA-256 on the host"},
                "modmul": {"executed_per_ciphertext": modmuls / (args.steps * max(1, n_local)),
                           "executed_frac_of_imad_peak": modmuls * macs / (ms_per_step * args.steps * 1e-3) / IMAD_PEAK_MAC_PER_S},
                "prover_s": prove_s, "proof_directory_bytes": nbytes}
        if world == 1 and not is_curve(args):
            # the same verification by the native pipeline (libvmnv.so: C++ over the C ABI, include/vmnv.h) from the
            # same host bytes -- a reported extra; a failure here never costs the line above
            try:
                vn = importlib.import_module("verificatum-vmn_b200.vmnv_native")
                VN = vn.MixNetElGamalVerifyFiatShamirSessionNative(G, params, 3, 2)
                if not VN.verify(nizkp)["accepted"]:
                    raise RuntimeError("the native verifier rejected an honest mix")
                torch.cuda.synchronize()
                t0 = time.time()
                for _ in range(args.steps):
                    VN.verify(nizkp)
                torch.cuda.synchronize()
                dt = (time.time() - t0) / args.steps
                line["native_verifier"] = {"value": n / dt, "unit": "ciphertexts/s", "ms_per_step": dt * 1e3,
                                           "gpu_launches_per_step": VN.report["launches"],
                                           "hashed_bytes_per_step": VN.report["hashed_bytes"],
                                           "what": "libvmnv.so: the whole verification in C++ over include/vmx.h, "
                                                   "from the proof directory's bytes in host memory (wall clock)"}
            except Exception as ex:
                line["native_verifier"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        if world == 1 and not args.no_cpu:
            try:
                from oracle import cpu_baseline
                res = cpu_baseline.run_verify_mix(bits=args.bits, n_total=n, sample=cpu_sample_size(args), group=args.group)
                line["cpu_baseline"] = {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"],
                                        "kind": "port", "sample": res["sample"]}
            except Exception as ex:  # a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "ciphertexts/s", "cores": 0, "kind": "port",
                                        "sample": "failed: %s" % ex}
    del V, M, w
    return line


def run_committed_shuffle(args, env):
    """BASELINE.json config 4's protocol: "re-encryption and expProd-heavy CCPoS verify" over multi-block
    ciphertexts.  After a pre-computation for N ciphertexts (permutation commitment + its proof of a shuffle of
    commitments; untimed, reported as `precomp_s`), one step = re-encryption factors pk^s for fresh exponents
    (2 omega fixed-base exponentiations per ciphertext), re-encryption + permutation, the commitment-consistent proof of
    a shuffle (hvzk/CCPoSW.java:75-150) and its verification (:160-260; mixnet/ShufflerElGamalSession.java:771-960)
    through the session API: published byte trees out of the prover, imported with membership checks by the
    verifier, Fiat-Shamir hashing on the host -- so `value` and `e2e` are the same end-to-end measurement.  One GPU."""
    torch = env.torch
    if env.world != 1:
        raise SystemExit("bench: --workload committed-shuffle runs on one GPU")
    vmx = importlib.import_module("verificatum-vmn_b200")
    A = vmx.arithm
    mixnet = importlib.import_module("verificatum-vmn_b200.mixnet")
    G = make_group(args, env)
    stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx), device=torch.device("cuda", env.local_rank))
    n, width = args.n, args.width
    setup_rs = make_prg("setup")
    x = G.getPRing().randomElement(setup_rs, 100)
    basic_pk = A.PPGroup(G, 2).product(G.getg(), G.getg().exp(x))
    exponentsRing = mixnet.getPlainPGroup(G, width).getPRing()
    widePk = mixnet.getWidePublicKey(basic_pk, width)
    if width == 1:
        ciphertexts = mixnet.demoCiphertexts(basic_pk, n, setup_rs)
    else:
        r0 = exponentsRing.randomElementArray(n, setup_rs, 100)
        ciphertexts = widePk.exp(r0)
        r0.free()
    params = mixnet.SessionParams(pGroupString=group_label(args))
    prover = mixnet.ShufflerSession(G, basic_pk, params, make_prg("prover"))
    verifier = mixnet.ShufflerSession(G, basic_pk, params, None)
    G.membership_check = True
    # ---- pre-computation (ShufflerElGamalSession.precomp :534-672) and its verification, untimed
    t0 = time.time()
    cs = mixnet.CommittedShuffler(prover, width, n)
    pub = cs.precomp()
    gens = verifier.deriveGenerators(n)
    pcv = mixnet.PermutationCommitment(verifier, gens)
    if not pcv.verify(*pub):
        raise SystemExit("bench: the proof of a shuffle of commitments was rejected")
    keep = cs.shrink(n)
    pcv.shrink(n, keep)
    G.sync()
    precomp_s = time.time() - t0

    def step(i: int):
        rs = make_prg("step%d" % i)
        prover.randomSource = rs
        for a in (cs.reencExponents, cs.reencFactors):
            a.free()
        cs.reencExponents = exponentsRing.randomElementArray(n, rs, params.rbitlen)
        cs.reencFactors = widePk.exp(cs.reencExponents)
        if args.offline_verify:
            proof, _ = cs.shuffle(ciphertexts)
            ok, out = mixnet.verifyCommittedShuffle(verifier, width, gens, pcv.commitment, ciphertexts, proof)
        else:   # the verifier reads the output from the board when it is published and hashes beside the prover
            ov = mixnet.OnlineCommittedVerification(verifier, width, gens, pcv.commitment, ciphertexts)
            proof, _ = cs.shuffle(ciphertexts, publish=ov.publish)
            ok, out = ov.finish(proof)
        out.free()
        if not ok:
            raise SystemExit("bench: the verifier rejected an honest commitment-consistent proof of a shuffle")
        return len(proof.output) + len(proof.commitment) + len(proof.reply)

    for i in range(max(1, args.warmup)):
        nbytes = step(i)
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    launches0, modmuls0 = G.launch_count(), G.modmul_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    t0 = time.time()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    e1.synchronize()
    wall = time.time() - t0
    sampler.stop_flag.set()
    sampler.join()
    ms_per_step = e0.elapsed_time(e1) / args.steps
    macs = 136 if is_curve(args) else macs_per_modmul(args.bits)
    modmuls = G.modmul_count() - modmuls0
    return {"metric": "ciphertexts/s: re-encrypt + CCPoS prove+verify after pre-computation, %s, width %d"
                      % (group_label(args), width),
            "value": n / (ms_per_step * 1e-3), "unit": "ciphertexts/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 limbs (exact integer)", "data": "synthetic",
            "config": {"workload": "%s, width %d, N=%d ciphertexts: re-encryption + commitment-consistent proof of a "
                                   "shuffle, prove + verify, after a pre-computation for N (BASELINE.json config 4's protocol)"
                                   % (group_label(args), width, n), "n_total": n, "width": width,
                       "cpu_sample": cpu_sample_size(args)},
            "clocks": sampler.summary(), "gpu_launches": int(G.launch_count() - launches0),
            "e2e": {"value": n / (wall / args.steps), "unit": "ciphertexts/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": nbytes, "ms_per_step": wall / args.steps * 1e3,
                    "includes": "byte-tree encode/decode, D2H/H2D, membership checks, Fiat-Shamir SHA-256 on the host",
                    "verifier": "offline" if args.offline_verify else
                                "online: the verifier hashes the output as the prover publishes it"},
            "modmul": {"executed_per_ciphertext": modmuls / (args.steps * n),
                       "executed_frac_of_imad_peak": modmuls * macs / (ms_per_step * args.steps * 1e-3) / IMAD_PEAK_MAC_PER_S},
            "precomp_s": precomp_s, "cpu_baseline": _cpu_committed(args)}


def _cpu_committed(args):
    if args.no_cpu or is_curve(args):
        return None
    try:
        from oracle import cpu_baseline
        res = cpu_baseline.run_committed_shuffle(bits=args.bits, n_total=args.n, sample=cpu_sample_size(args),
                                                 group=args.group, width=args.width)
        return {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"], "kind": "port",
                "sample": res["sample"]}
    except Exception as ex:  # a reported number, never a reason to lose the GPU line
        return {"value": None, "unit": "ciphertexts/s", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def run_shuffle(args, env):
    """One list of N ciphertexts: re-encrypt + permute, PoSBasicTW prove, PoSBasicTW verify (BASELINE.json's metric;
    `configs[1]` at --n 100000).  Returns the JSON line as a dict on rank 0, None elsewhere."""
    import numpy as np
    torch, world, rank, local_rank = env.torch, env.world, env.rank, env.local_rank
    vmx = importlib.import_module("verificatum-vmn_b200")
    A = vmx.arithm
    hvzk = importlib.import_module("verificatum-vmn_b200.hvzk")
    mixnet = importlib.import_module("verificatum-vmn_b200.mixnet")
    crypto = vmx.crypto
    prg = make_prg

    G = make_group(args, env)
    stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx), device=torch.device("cuda", local_rank))
    n = args.n                                              # global list size
    n_local = n * (rank + 1) // world - n * rank // world   # this rank's contiguous index range

    # ---- synthetic inputs, resident in HBM
    setup_rs = prg("setup")
    x = G.getPRing().randomElement(setup_rs, 100)
    y = G.getg().exp(x)
    pk = A.PPGroup(G, 2).product(G.getg(), y)
    width = args.width
    exponentsRing = mixnet.getPlainPGroup(G, width).getPRing()
    if width == 1:
        ciphertexts = mixnet.demoCiphertexts(pk, n, setup_rs)
    else:  # multi-block ciphertexts: w = widePk^r, r in the product ring (encryptions of the unit element)
        r0 = exponentsRing.randomElementArray(n, setup_rs, 100)
        ciphertexts = mixnet.getWidePublicKey(pk, width).exp(r0)
        r0.free()
    basic_pk, pk = pk, mixnet.getWidePublicKey(pk, width)
    params = mixnet.SessionParams(pGroupString="ECqPGroup(%s)" % args.group if is_curve(args)
                                  else "ModPGroup(RFC3526-%d)" % args.bits)
    session = mixnet.ShufflerSession(G, basic_pk, params, prg("prover"))
    generators = session.deriveGenerators(n)
    seed = bytes(range(32))
    challenge = int.from_bytes(crypto.HashfunctionHeuristic("SHA-256").hash(b"challenge"), "big")
    G.sync()

    phase_ms = {}

    class Phase:
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(stream)

        def __exit__(self, *a):
            self.e1.record(stream)
            self.e1.synchronize()
            phase_ms[self.name] = phase_ms.get(self.name, 0.0) + self.e0.elapsed_time(self.e1)

    def step_device(i: int, timed: bool):
        """Re-encrypt + prove + verify with everything resident on the device."""
        rs = prg("step%d" % i)
        P = hvzk.PoSBasicTW(params.vbitlenro, params.ebitlenro, params.rbitlen, crypto.PRGHeuristic(), rs)
        V = hvzk.PoSBasicTW(params.vbitlenro, params.ebitlenro, params.rbitlen, crypto.PRGHeuristic(), rs)
        ph = Phase if (timed and args.phases) else (lambda name: _Null())
        with ph("reencrypt"):
            s = exponentsRing.randomElementArray(n, rs, params.rbitlen)
            factors = pk.exp(s)
            pi = A.Permutation.random(n, rs, params.rbitlen, G)
            reenc = ciphertexts.mul(factors)
            factors.free()
            inv = pi.inv()
            out = reenc.permute(inv)
            reenc.free()
        with ph("prove.precompute"):
            P.precompute(G.getg(), generators, pi)
        with ph("prove.commit"):
            P.setInstance(pk, ciphertexts, out, s)
            P.commit(seed)
        with ph("prove.reply"):
            P.reply(challenge)
        with ph("verify.computeAF"):
            V.precompute(G.getg(), generators)
            V.setInstance(pk, ciphertexts, out)
            V.u = P.u
            V.setBatchVector(seed)
            V.computeAF()
        with ph("verify.checks"):
            V.B, V.Ap, V.Bp, V.Cp, V.Dp, V.Fp = P.B, P.Ap, P.Bp, P.Cp, P.Dp, P.Fp
            V.setChallenge(challenge)
            V.k_A, V.k_B, V.k_C, V.k_D, V.k_E, V.k_F = P.k_A, P.k_B, P.k_C, P.k_D, P.k_E, P.k_F
            ok = V.verifyParsed()
        if not ok:
            raise SystemExit("bench: the verifier rejected an honest proof: %r" % (V.verdicts,))
        V.e.free()
        V.u = V.e = V.B = V.Bp = V.k_B = V.k_E = None
        P.free()
        s.free()
        out.free()

    # ---- the first step of a session builds the fixed-base tables of g, the public key and h0: timed on its own
    env.barrier()
    c0 = torch.cuda.Event(enable_timing=True)
    c1 = torch.cuda.Event(enable_timing=True)
    mm_cold0 = G.modmul_count()
    c0.record(stream)
    step_device(0, False)
    c1.record(stream)
    c1.synchronize()
    cold_ms = env.max_over_ranks(c0.elapsed_time(c1))
    cold_modmuls = G.modmul_count() - mm_cold0

    # ---- device-resident timing
    for i in range(1, args.warmup):
        step_device(i, False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    env.barrier()
    launches0, modmuls0 = G.launch_count(), G.modmul_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    e0.record(stream)
    marks = []
    for i in range(args.steps):
        step_device(args.warmup + i, True)
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record(stream)
    e1.record(stream)
    e1.synchronize()
    env.barrier()
    t_host = time.time() - t_host0
    step_ms = [(marks[i - 1] if i else e0).elapsed_time(marks[i]) for i in range(len(marks))]
    launches = G.launch_count() - launches0
    modmuls = G.modmul_count() - modmuls0
    sampler.stop_flag.set()
    sampler.join()
    dev_ms = env.max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = dev_ms / args.steps
    value = n / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (fixed-base exponentiation), timed live
    # every rank runs it on its shard (no collective inside); rank 0 reports its own
    rs = prg("roofline")
    e = G.getPRing().randomElementArray(n, rs, params.rbitlen)
    tmp = G.getg().exp(e)
    tmp.free()
    mm0 = G.modmul_count()
    r0 = torch.cuda.Event(enable_timing=True)
    r1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    r0.record(stream)
    for _ in range(reps):
        tmp = G.getg().exp(e)
        tmp.free()
    r1.record(stream)
    r1.synchronize()
    k_ms = r0.elapsed_time(r1) / reps
    k_modmuls = (G.modmul_count() - mm0) / reps
    e.free()
    macs = 136 if is_curve(args) else macs_per_modmul(args.bits)
    elem_bytes = 64 if is_curve(args) else args.bits // 8
    achieved = k_modmuls * macs / (k_ms * 1e-3)
    traffic = pipe_busy = traffic_src = None
    per_exp = k_modmuls / max(1, n_local)
    try:  # an `ncu --set full` capture of this kernel counts only if it ran the SAME launch (size and window)
        for name in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
            if not name.endswith(".json") or "ncu_exp_fixed" not in name:
                continue
            cap = json.load(open(os.path.join(ROOT, "profiles", name)))
            if cap.get("n") == n_local and cap.get("bits", 3072) == args.bits and not is_curve(args) and \
                    abs(cap.get("modmuls_per_exponent", -1) - per_exp) < 0.5:
                traffic, pipe_busy, traffic_src = cap["traffic_bytes"], cap["fmaheavy_pipe_busy_pct"], "profiles/" + name
    except Exception:
        pass
    roof = {"bound": "imad", "kernel": "k_ec_exp_fixed" if is_curve(args) else "k_exp_fixed<%d>" % (args.bits // 32),
            "achieved": achieved / 1e12,
            "peak": IMAD_PEAK_MAC_PER_S / 1e12, "unit": "TMAC/s (32x32+64 IMAD.WIDE)", "frac": achieved / IMAD_PEAK_MAC_PER_S,
            "traffic": traffic,
            "traffic_unit": ("bytes of DRAM read+write per launch, ncu --set full of the same launch (%s)" % traffic_src)
            if traffic_src else "no ncu capture of this exact launch (size, window) is committed: null",
            "algorithmic_bytes": (k_modmuls / 11 * elem_bytes + 32 * n_local + 96 * n_local) if is_curve(args)
            else k_modmuls * elem_bytes + 2 * n_local * elem_bytes,
            "unit_of_work": "field multiplication = 136 word MACs nominal (the P-256 reduction executes 64 + adds)"
            if is_curve(args) else "modmul = 2N^2+N word MACs",
            "imad_pipe_busy_pct_ncu": pipe_busy, "modmuls_per_launch": k_modmuls, "modmuls_per_exponent": per_exp,
            "ms_per_launch": k_ms,
            "peak_source": "measured on B200 (profiles/r01_ubench_imad.txt); MEASURED_PEAKS.json has no integer peak",
            "hbm_note": "integer-pipe bound: arithmetic intensity ~1e4 MAC/B, HBM is not the limiter"}

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        G.membership_check = os.environ.get("VMX_BENCH_MEMBERSHIP", "1") == "1"
        ciph_bytes = ciphertexts.toByteTree().to_bytes()
        pinned = torch.empty(len(ciph_bytes), dtype=torch.uint8).pin_memory()
        pinned.numpy()[:] = np.frombuffer(ciph_bytes, dtype=np.uint8)
        del ciph_bytes
        h2d = d2h = hashed = 0

        tracemod = importlib.import_module("verificatum-vmn_b200._trace")
        _span = tracemod.span

        def step_e2e(i: int):
            nonlocal h2d, d2h
            prover = mixnet.ShufflerSession(G, basic_pk, params, prg("e2e%d" % i))
            verifier = mixnet.ShufflerSession(G, basic_pk, params, prg("e2ev%d" % i))
            ciphPGroup = mixnet.getCiphPGroup(G, width)
            w = ciphPGroup.toElementArray(n, vmx.eio.ByteTreeReader(memoryview(pinned.numpy()).toreadonly()))
            if args.offline_verify:
                with _span("e2e.shuffle"):
                    proof, _ = prover.shuffle(width, w, generators=generators)
                with _span("e2e.verify"):
                    ok, out = verifier.verify(width, w, proof, generators=generators)
            else:
                # the verifier follows the bulletin board, as the reference's mix-servers do (hvzk/PoSTW.java:195-245):
                # it hashes each message as the prover publishes it, then verifies
                ov = verifier.beginVerify(width, w, generators=generators)
                with _span("e2e.shuffle"):
                    proof, _ = prover.shuffle(width, w, generators=generators, publish=ov.publish)
                with _span("e2e.verify"):
                    ok, out = ov.finish(proof)
            if not ok:
                raise SystemExit("bench e2e: verifier rejected an honest proof")
            out.free()
            w.free()
            h2d = pinned.numel() + len(proof.output) + len(proof.permutationCommitment) + len(proof.commitment) + \
                len(proof.reply)
            d2h = len(proof.output) + len(proof.permutationCommitment) + len(proof.commitment) + len(proof.reply)

        step_e2e(0)
        if args.phases and rank == 0 and world == 1:  # where does the host time go?  (untimed extra step)
            import cProfile
            import pstats
            import io
            pr = cProfile.Profile()
            pr.enable()
            step_e2e(99)
            pr.disable()
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(18)
            sys.stderr.write(buf.getvalue())
        if args.trace:   # every rank runs the step (it is full of collectives); rank 0 keeps its timeline
            tracemod.start()
            with tracemod.span("e2e.step"):
                step_e2e(98)
            events = tracemod.stop()
            if rank == 0:
                with open(args.trace, "w") as f:
                    json.dump(events, f)
        hash0 = crypto.hashed_bytes() if hasattr(crypto, "hashed_bytes") else None
        hsec0 = crypto.hashed_seconds() if hasattr(crypto, "hashed_seconds") else None
        env.barrier()
        t0 = time.time()
        k = args.e2e_steps
        for i in range(k):
            step_e2e(1 + i)
        env.barrier()
        dt = env.max_over_ranks((time.time() - t0) / k)
        e2e = {"value": n / dt, "unit": "ciphertexts/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": k,
               "includes": "byte-tree decode/encode, H2D/D2H, Fiat-Shamir SHA-256 on the host",
               "membership_check_on_import": bool(G.membership_check),
               "verifier": "offline: verify(proof) after shuffle() returned" if args.offline_verify else
               "online: the verifier hashes each message as the prover publishes it (bulletin-board order), "
               "verdict after the last message"}
        if hash0 is not None:
            hb = (crypto.hashed_bytes() - hash0) / k
            hs = (crypto.hashed_seconds() - hsec0) / k
            e2e["hash_bytes_per_step"] = hb
            e2e["sha256_gbs"] = hb / hs / 1e9 if hs > 0 else None
            e2e["sha256_bound_ciphertexts_per_s"] = n / hs if hs > 0 else None
            e2e["hash_note"] = ("Fiat-Shamir is ONE SHA-256 stream per challenge (hvzk/ChallengerRO.java:96-116): "
                                "hash_bytes_per_step at sha256_gbs is the Amdahl bound of the end-to-end figure at any "
                                "number of GPUs (rank 0's hashing)")

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.width == 1:
        try:
            from oracle import cpu_baseline
            res = cpu_baseline.run(bits=args.bits, n_total=n, sample=cpu_sample_size(args), steps=1, warmup=0,
                                   group=args.group)
            cpu = {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"], "kind": "port",
                   "sample": res["sample"]}
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the GPU line
            cpu = {"value": None, "unit": "ciphertexts/s", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}

    line = None
    if rank == 0:
        nominal = nominal_fieldmuls_per_ciphertext_ec(n, width=args.width) if is_curve(args) else \
            nominal_modmuls_per_ciphertext(args.bits, n, width=args.width)
        line = {"metric": metric_name(args),
                "value": value, "unit": "ciphertexts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32 limbs (exact integer)", "data": "synthetic", "config": config_dict(args, world),
                "clocks": sampler.summary(), "gpu_launches": int(launches), "e2e": e2e, "roofline": roof,
                "cpu_baseline": cpu,
                "cold_first_step_ms": cold_ms,
                "cold_first_step_note": "the first step of a session, fixed-base tables of g, pk and h0 built inside "
                                        "(%d modmuls against %d of a warm step); a mix-server builds them once per "
                                        "session" % (cold_modmuls, modmuls // max(1, args.steps)),
                "modmul": {"nominal_per_ciphertext": nominal["total"],
                           "executed_per_ciphertext": modmuls / (args.steps * max(1, n_local)),
                           "nominal_modmul_per_s": value * nominal["total"],
                           "nominal_frac_of_imad_peak": value / world * nominal["total"] * macs / IMAD_PEAK_MAC_PER_S,
                           "executed_frac_of_imad_peak": modmuls * macs / (dev_ms * 1e-3) / IMAD_PEAK_MAC_PER_S,
                           "note": "fractions are per GPU (rank 0's kernels against one GPU's peak); a squaring "
                                   "counts as one modmul although it executes ~0.83 of a multiplication's MACs"},
                "host_wall_ms_per_step": t_host * 1e3 / args.steps,
                "step_ms": {"min": min(step_ms), "max": max(step_ms), "all": step_ms[:32]}}
        if args.phases:
            line["phase_ms_per_step"] = {k: v / args.steps for k, v in phase_ms.items()}
        try:  # how this line relates to the headline metric of BASELINE.json
            with open(os.path.join(ROOT, "BASELINE.json")) as f:
                line["baseline_metric"] = {"metric": json.load(f)["metric"],
                                           "relation": "same unit, path and list size when --n is 1000000 (the default); "
                                                       "one step here also includes the PROVER (re-encryption + "
                                                       "proof-of-shuffle prove AND verify)"}
        except Exception:
            pass
    ciphertexts.free()
    generators.free()
    del G
    return line


# BASELINE.json configs 3, 4, 5 at sizes whose set-up fits a side run (the mix of config 3 at N = 10^6 takes the
# three decryption servers ~30 s each to produce; config 5 at its own N = 10^7): short runs in the same process,
# after the headline measurement
OTHER_CONFIGS = [
    ("config3_verify_mix_3072", dict(workload="verify-mix", bits=3072, group="modp", width=1, n=100000)),
    ("config4_width3_2048", dict(workload="shuffle", bits=2048, group="modp", width=3, n=100000)),
    ("config5_p256", dict(workload="shuffle", bits=3072, group="P-256", width=1, n=10000000)),
]


def run_isolated(extra, timeout_s: float):
    """A side run in its own process (own CUDA context): whatever happens to it cannot cost the headline line."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--no-other", "--no-cpu", "--steps", "2", "--warmup", "2"] + extra
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    t0 = time.time()
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
    except subprocess.TimeoutExpired:
        return {"error": "no result within %d s" % timeout_s}
    for ln in reversed(p.stdout.splitlines()):
        if ln.startswith("{"):
            try:
                line = json.loads(ln)
                line["wall_s"] = time.time() - t0
                return line
            except ValueError:
                break
    return {"error": "rc=%d: %s" % (p.returncode, (p.stderr or p.stdout)[-300:])}


def run_other_configs(args, env):
    import copy
    import gc
    out = {}
    for name, over in OTHER_CONFIGS:
        # beyond one GPU only config 3 (the one BASELINE.json defines "sharded over 1/2/4/8 B200"): a side run that
        # fails on one rank would leave the others waiting in a collective
        if env.world > 1 and not name.startswith("config3"):
            continue
        a = copy.copy(args)
        for k, v in over.items():
            setattr(a, k, v)
        a.steps, a.warmup, a.e2e_steps = 2, 2, 2
        a.no_cpu, a.phases, a.trace, a.cpu_sample = True, False, "", 0
        gc.collect()
        t0 = time.time()
        try:
            line = run_verify_mix(a, env) if a.workload == "verify-mix" else run_shuffle(a, env)
        except SystemExit as ex:
            line = {"error": str(ex)}
        except Exception as ex:  # a side measurement never costs the headline line
            line = {"error": "%s: %s" % (type(ex).__name__, ex)}
        if env.rank == 0 and line is not None:
            keep = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "n_gpus", "scaling", "gpu_launches",
                    "e2e", "modmul", "cold_first_step_ms", "native_verifier", "error")
            sub = {k: line[k] for k in keep if k in line}
            if "config" in line:
                sub["workload"] = line["config"]["workload"]
            if "roofline" in line and line["roofline"]:
                sub["roofline"] = {k: line["roofline"][k] for k in ("kernel", "achieved", "peak", "unit", "frac")}
            sub["wall_s"] = time.time() - t0
            out[name] = sub
    if env.world == 1:
        # config 4's own protocol (pre-computation, then re-encryption + commitment-consistent proof of a shuffle),
        # in a process of its own
        gc.collect()
        line = run_isolated(["--workload", "committed-shuffle", "--bits", "2048", "--width", "3", "--n", "100000"], 240)
        keep = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "gpu_launches", "e2e", "modmul", "precomp_s",
                "wall_s", "error")
        out["config4_ccpos_2048_width3"] = {k: line[k] for k in keep if k in line}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vmx", choices=["vmx", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("VMX_BENCH_N", "1000000")),
                    help="ciphertexts in the list, in TOTAL (sharded over the GPUs): BASELINE.json's metric is N = 10^6")
    ap.add_argument("--bits", type=int, default=3072)
    ap.add_argument("--width", type=int, default=1, help="blocks per ciphertext (BASELINE.json config 4 uses 3)")
    ap.add_argument("--group", default="modp", help="modp (RFC 3526 safe prime of --bits) or a curve name (P-256)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="ciphertexts in the CPU baseline sample (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=5, help="end-to-end steps averaged (after one untimed)")
    ap.add_argument("--offline-verify", action="store_true",
                    help="end to end: verify only after the whole proof was returned (default: the verifier "
                         "follows the messages as they are published)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the short side runs of BASELINE.json configs 3, 4, 5")
    ap.add_argument("--phases", action="store_true", help="print per-phase device times to stderr")
    ap.add_argument("--trace", default="", help="write the host timeline (ABI calls, hashing) of one extra untimed "
                                                "end-to-end step to this JSON file")
    ap.add_argument("--workload", default="shuffle", choices=["shuffle", "verify-mix", "committed-shuffle"],
                    help="shuffle: re-encrypt + PoS prove + verify (BASELINE.json's metric, the default); "
                         "verify-mix: vmnv-style verification of a 3-party mix, threshold 2 (config 3)")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)
    if args.trace:
        os.environ["VMX_TRACE"] = "1"

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = Env(torch, dist, world, rank, local_rank)

    headline = (args.workload == "shuffle" and args.group == "modp" and args.bits == 3072 and args.width == 1)
    if args.workload == "committed-shuffle":
        line = run_committed_shuffle(args, env)
    else:
        line = run_verify_mix(args, env) if args.workload == "verify-mix" else run_shuffle(args, env)
    if headline and not args.no_other:
        other = run_other_configs(args, env)
        if rank == 0:
            line["other_configs"] = other
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
