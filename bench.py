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

# measured on B200 by scratch/ubench.cu (profiles/r01_ubench_imad.txt): IMAD.WIDE.U32 issues at
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
    if args.workload == "verify-mix":
        res = cpu_baseline.run_verify_mix(bits=args.bits, n_total=args.n, sample=args.cpu_sample, steps=args.steps,
                                          warmup=min(args.warmup, 1), group=args.group)
    else:
        res = cpu_baseline.run(bits=args.bits, n_total=args.n, sample=args.cpu_sample, steps=args.steps,
                               warmup=min(args.warmup, 1), group=args.group)
    mixw = args.workload == "verify-mix"
    line = {"metric": mix_metric_name(args) if mixw else metric_name(args), "impl": "reference",
            "value": res["value"], "unit": "ciphertexts/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (exact integer)", "data": "synthetic",
            "config": mix_config_dict(args) if mixw else config_dict(args),
            "cpu_baseline": {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"], "kind": "port",
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "ciphertexts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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


def mix_config_dict(args):
    return dict(config_dict(args), workload="%s, width 1, N=%d ciphertexts per GPU: verification of a 3-party "
                "mix with threshold 2 from its proof directory in host memory" % (group_label(args), args.n),
                k=3, threshold=2)


def config_dict(args):
    elem = 64 if is_curve(args) else args.bits // 8
    return {"workload": "%s, width %d, N=%d ciphertexts per GPU: re-encrypt + PoSBasicTW prove + verify"
                        % (group_label(args), args.width, args.n),
            "group": args.group, "bits": 256 if is_curve(args) else args.bits, "width": args.width, "n_per_gpu": args.n,
            "ebitlen": 256, "vbitlen": 256, "rbitlen": 100,
            "l2": "working set (N x %d B per array, >10 arrays) exceeds the 126 MB L2" % elem}


def run_verify_mix(args, G, prg, stream, world, rank, local_rank, torch, dist):
    """BASELINE.json config 3: what `vmnv` does with the proof directory of a 3-party mix (threshold 2): two
    verifyPoS and the verification of the decryption (3 arrays of decryption factors, batched proof, plaintexts)
    -- mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668.  The proof directory is produced by the engine's
    own mix (untimed) and held in HOST memory; one step = one full verification from those bytes (import with
    membership checks, Fiat-Shamir hashing, all array work), so `value` and `e2e` are the same measurement here."""
    vm = importlib.import_module("verificatum-vmn_b200.vmnv")
    mixnet = importlib.import_module("verificatum-vmn_b200.mixnet")
    n = args.n * world
    params = mixnet.SessionParams(pGroupString="%s" % group_label(args))
    M = vm.MixNetElGamal(G, params, 3, 2, prg("mix/dealer"))
    w = mixnet.demoCiphertexts(M.fullPublicKey, n, prg("mix/input"))
    t0 = time.time()
    M.run(w).free()
    G.sync()
    prove_s = time.time() - t0
    nizkp = M.nizkp
    V = vm.MixNetElGamalVerifyFiatShamirSession(G, params, 3, 2)
    G.membership_check = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        if not V.verify(nizkp)["accepted"]:
            raise SystemExit("bench: the verifier rejected an honest mix")
    if args.trace and rank == 0:
        tr = importlib.import_module("verificatum-vmn_b200._trace")
        tr.start()
        with tr.span("e2e.step"):
            V.verify(nizkp)
        with open(args.trace, "w") as f:
            json.dump(tr.stop(), f)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0, modmuls0 = G.launch_count(), G.modmul_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.time()
    for _ in range(args.steps):
        V.verify(nizkp)
    e1.record(stream)
    e1.synchronize()
    barrier()
    wall = time.time() - t0
    sampler.stop_flag.set()
    sampler.join()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    if rank == 0:
        macs = 136 if is_curve(args) else macs_per_modmul(args.bits)
        nbytes = sum(len(v) for v in nizkp.values())
        modmuls = G.modmul_count() - modmuls0
        line = {"metric": mix_metric_name(args),
                "value": n / (ms_per_step * 1e-3), "unit": "ciphertexts/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32 limbs (exact integer)", "data": "synthetic",
                "config": mix_config_dict(args),
                "clocks": sampler.summary(), "gpu_launches": int(G.launch_count() - launches0),
                "e2e": {"value": n / (wall / args.steps), "unit": "ciphertexts/s", "h2d_bytes_per_step": nbytes,
                        "d2h_bytes_per_step": 0, "ms_per_step": wall / args.steps * 1e3,
                        "includes": "byte-tree decode, H2D, membership checks, Fiat-Shamir SHA-256 on the host"},
                "modmul": {"executed_per_ciphertext": modmuls / (args.steps * args.n),
                           "executed_frac_of_imad_peak": modmuls * macs / (ms_per_step * args.steps * 1e-3) / IMAD_PEAK_MAC_PER_S},
                "prover_s": prove_s, "proof_directory_bytes": nbytes}
        if world == 1 and not args.no_cpu:
            try:
                from oracle import cpu_baseline
                res = cpu_baseline.run_verify_mix(bits=args.bits, n_total=n, sample=args.cpu_sample, group=args.group)
                line["cpu_baseline"] = {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"],
                                        "kind": "port", "sample": res["sample"]}
            except Exception as ex:  # a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "ciphertexts/s", "cores": 0, "kind": "port",
                                        "sample": "failed: %s" % ex}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vmx", choices=["vmx", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("VMX_BENCH_N", "100000")))
    ap.add_argument("--bits", type=int, default=3072)
    ap.add_argument("--width", type=int, default=1, help="blocks per ciphertext (BASELINE.json config 4 uses 3)")
    ap.add_argument("--group", default="modp", help="modp (RFC 3526 safe prime of --bits) or a curve name (P-256)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="ciphertexts in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--phases", action="store_true", help="print per-phase device times to stderr")
    ap.add_argument("--trace", default="", help="write the host timeline (ABI calls, hashing) of one extra untimed "
                                                "end-to-end step to this JSON file")
    ap.add_argument("--workload", default="shuffle", choices=["shuffle", "verify-mix"],
                    help="shuffle: re-encrypt + PoS prove + verify (BASELINE.json config 2, the default); "
                         "verify-mix: vmnv-style verification of a 3-party mix, threshold 2 (config 3)")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)
    if args.trace:
        os.environ["VMX_TRACE"] = "1"

    import numpy as np
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

    vmx = importlib.import_module("verificatum-vmn_b200")
    A = vmx.arithm
    hvzk = importlib.import_module("verificatum-vmn_b200.hvzk")
    mixnet = importlib.import_module("verificatum-vmn_b200.mixnet")
    groups = importlib.import_module("verificatum-vmn_b200.groups")
    crypto = vmx.crypto

    p, q, g = groups.rfc3526(args.bits)
    if is_curve(args):
        if world > 1:
            par = importlib.import_module("verificatum-vmn_b200.parallel")
            G = par.make_curve_group(args.group, local_rank)
        else:
            G = A.ECqPGroup(args.group, device=local_rank)
    elif world > 1:
        # ONE list of world * n ciphertexts, sharded in contiguous index ranges over the GPUs: one
        # shuffle, one proof; expProd partial products and permuted rows travel over NCCL
        par = importlib.import_module("verificatum-vmn_b200.parallel")
        G = par.make_group(p, q, g, local_rank)
    else:
        G = A.ModPGroup(p, q, g, device=local_rank)
    stream = torch.cuda.ExternalStream(G._lib.vmx_ctx_stream(G.ctx), device=torch.device("cuda", local_rank))
    n_local = args.n
    n = args.n * world          # global list size (every rank holds n_local of every array)

    def prg(label: str):
        # the same stream on every rank: each rank expands its own slice of it on the device
        r = crypto.PRGHeuristic()
        r.setSeed(crypto.HashfunctionHeuristic("SHA-256").hash(("vmx-bench/%s" % label).encode()))
        return r

    if args.workload == "verify-mix":
        return run_verify_mix(args, G, prg, stream, world, rank, local_rank, torch, dist)

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

    class _Null:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for i in range(args.warmup):
        step_device(i, False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0, modmuls0 = G.launch_count(), G.modmul_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i, True)
    e1.record(stream)
    e1.synchronize()
    barrier()
    t_host = time.time() - t_host0
    dev_ms = e0.elapsed_time(e1)
    launches = G.launch_count() - launches0
    modmuls = G.modmul_count() - modmuls0
    sampler.stop_flag.set()
    sampler.join()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = n / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (fixed-base exponentiation), timed live
    roof = None
    if True:  # every rank runs it on its shard (no collective inside); rank 0 reports its own
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
        traffic = pipe_busy = None
        try:  # one `ncu --set full` capture of this kernel at the same shape (profiles/, per launch)
            cap = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_exp_fixed.json")))
            if cap["n"] == n_local and args.bits == 3072 and not is_curve(args):
                traffic, pipe_busy = cap["traffic_bytes"], cap["fmaheavy_pipe_busy_pct"]
        except Exception:
            pass
        roof = {"bound": "imad", "kernel": "k_ec_exp_fixed" if is_curve(args) else "k_exp_fixed<%d>" % (args.bits // 32),
                "achieved": achieved / 1e12,
                "peak": IMAD_PEAK_MAC_PER_S / 1e12, "unit": "TMAC/s (32x32+64 IMAD.WIDE)", "frac": achieved / IMAD_PEAK_MAC_PER_S,
                "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch (ncu, profiles/r02_ncu_exp_fixed.json; 16-byte gathers use half of each 32-byte sector, hence ~2x the algorithmic bytes; 0.34 TB/s, not the limiter)",
                "algorithmic_bytes": (k_modmuls / 11 * elem_bytes + 32 * n_local + 96 * n_local) if is_curve(args)
                else k_modmuls * elem_bytes + 2 * n_local * elem_bytes,
                "unit_of_work": "field multiplication = 136 word MACs nominal (the P-256 reduction executes 64 + adds)"
                if is_curve(args) else "modmul = 2N^2+N word MACs",
                "imad_pipe_busy_pct_ncu": pipe_busy, "modmuls_per_launch": k_modmuls, "ms_per_launch": k_ms,
                "peak_source": "measured on B200 (profiles/r01_ubench_imad.txt); MEASURED_PEAKS.json has no integer peak",
                "hbm_note": "integer-pipe bound: arithmetic intensity ~1e4 MAC/B, HBM is not the limiter"}

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        G.membership_check = os.environ.get("VMX_BENCH_MEMBERSHIP", "1") == "1"
        ciph_bytes = ciphertexts.toByteTree().to_bytes()
        pinned = torch.empty(len(ciph_bytes), dtype=torch.uint8).pin_memory()
        pinned.numpy()[:] = np.frombuffer(ciph_bytes, dtype=np.uint8)
        h2d = d2h = 0

        _span = importlib.import_module("verificatum-vmn_b200._trace").span

        def step_e2e(i: int):
            nonlocal h2d, d2h
            prover = mixnet.ShufflerSession(G, basic_pk, params, prg("e2e%d" % i))
            verifier = mixnet.ShufflerSession(G, basic_pk, params, prg("e2ev%d" % i))
            ciphPGroup = mixnet.getCiphPGroup(G, width)
            w = ciphPGroup.toElementArray(n, vmx.eio.ByteTreeReader(memoryview(pinned.numpy()).toreadonly()))
            with _span("e2e.shuffle"):
                proof, _ = prover.shuffle(width, w, generators=generators)
            with _span("e2e.verify"):
                ok, out = verifier.verify(width, w, proof, generators=generators)
            if not ok:
                raise SystemExit("bench e2e: verifier rejected an honest proof")
            out.free()
            w.free()
            h2d = len(ciph_bytes) + len(proof.output) + len(proof.permutationCommitment) + len(proof.commitment) + \
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
        if args.trace and rank == 0:
            tr = importlib.import_module("verificatum-vmn_b200._trace")
            tr.start()
            with tr.span("e2e.step"):
                step_e2e(98)
            with open(args.trace, "w") as f:
                json.dump(tr.stop(), f)
        barrier()
        t0 = time.time()
        k = max(1, min(args.steps, 2))
        for i in range(k):
            step_e2e(1 + i)
        barrier()
        dt = (time.time() - t0) / k
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n / float(tt.item()), "unit": "ciphertexts/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(tt.item()) * 1e3,
               "includes": "byte-tree decode/encode, H2D/D2H, Fiat-Shamir SHA-256 on the host",
               "membership_check_on_import": bool(G.membership_check)}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.width == 1:
        try:
            from oracle import cpu_baseline
            res = cpu_baseline.run(bits=args.bits, n_total=n, sample=args.cpu_sample, steps=1, warmup=0,
                                   group=args.group)
            cpu = {"value": res["value"], "unit": "ciphertexts/s", "cores": res["cores"], "kind": "port",
                   "sample": res["sample"]}
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the GPU line
            cpu = {"value": None, "unit": "ciphertexts/s", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}

    if rank == 0:
        nominal = nominal_fieldmuls_per_ciphertext_ec(n, width=args.width) if is_curve(args) else \
            nominal_modmuls_per_ciphertext(args.bits, n, width=args.width)
        macs = 136 if is_curve(args) else macs_per_modmul(args.bits)
        line = {"metric": metric_name(args),
                "value": value, "unit": "ciphertexts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 limbs (exact integer)", "data": "synthetic", "config": config_dict(args),
                "clocks": sampler.summary(), "gpu_launches": int(launches), "e2e": e2e, "roofline": roof,
                "cpu_baseline": cpu,
                "modmul": {"nominal_per_ciphertext": nominal["total"],
                           "executed_per_ciphertext": modmuls / (args.steps * n_local),
                           "nominal_modmul_per_s": value * nominal["total"],
                           "nominal_frac_of_imad_peak": value / world * nominal["total"] * macs / IMAD_PEAK_MAC_PER_S,
                           "executed_frac_of_imad_peak": modmuls * macs / (dev_ms * 1e-3) / IMAD_PEAK_MAC_PER_S,
                           "note": "fractions are per GPU (rank 0's kernels against one GPU's peak)"},
                "host_wall_ms_per_step": t_host * 1e3 / args.steps}
        if args.phases:
            line["phase_ms_per_step"] = {k: v / args.steps for k, v in phase_ms.items()}
        try:  # how this line relates to the headline metric of BASELINE.json
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "BASELINE.json")) as f:
                line["baseline_metric"] = {"metric": json.load(f)["metric"],
                                           "relation": "same unit and path; one step here also includes the PROVER "
                                                       "(configs[1]: re-encryption + proof-of-shuffle prove/verify) and "
                                                       "runs at N = %d per GPU; --n 1000000 gives the metric's N" % args.n}
        except Exception:
            pass
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
