#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 170 python bench.py --steps 1 --warmup 2 --e2e-steps 3 --no-cpu --no-other > gpurun_out/s9_bench_1m.log 2> gpurun_out/s9_bench_1m.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s9_bench_1m.err
