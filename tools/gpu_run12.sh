#!/bin/bash
# round-2 GPU session 12 (1 GPU): config 4's protocol after the CCPoS commitment was moved beside the seed hash and the
# verifier made online; offline verification for comparison
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 40 python bench.py --workload committed-shuffle --bits 2048 --width 3 --n 100000 --steps 3 --warmup 1 --no-cpu --no-other > gpurun_out/s12_bench_ccpos_online.log 2> gpurun_out/s12_bench_ccpos_online.err; echo "bench rc=$?"
timeout 40 python bench.py --workload committed-shuffle --bits 2048 --width 3 --n 100000 --steps 3 --warmup 1 --no-cpu --no-other --offline-verify > gpurun_out/s12_bench_ccpos_offline.log 2> gpurun_out/s12_bench_ccpos_offline.err; echo "bench rc=$?"
for f in online offline; do python - <<P
import json
for l in open("gpurun_out/s12_bench_ccpos_$f.log"):
    if l.startswith("{"):
        d = json.loads(l); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["modmul"])
P
done
tail -c 300 gpurun_out/s12_bench_ccpos_online.err
