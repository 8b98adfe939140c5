#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 150 python bench.py --workload verify-mix --n 100000 --steps 3 --warmup 1 --no-cpu > gpurun_out/s8_bench_mix.log 2> gpurun_out/s8_bench_mix.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s8_bench_mix.err
