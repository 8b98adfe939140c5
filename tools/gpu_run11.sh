#!/bin/bash
# round-2 GPU session 11 (1 GPU): BASELINE config 4's protocol (pre-computation, then re-encryption + CCPoS prove + verify) at N = 10^5
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 85 python bench.py --workload committed-shuffle --bits 2048 --width 3 --n 100000 --steps 2 --warmup 1 --no-cpu --no-other > gpurun_out/s11_bench_ccpos.log 2> gpurun_out/s11_bench_ccpos.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/s11_bench_ccpos.log; tail -c 600 gpurun_out/s11_bench_ccpos.err
