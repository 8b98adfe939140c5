#!/bin/bash
# round-2 GPU session 6 (1 GPU): the default command as the driver runs it (fewer steps), incl. other_configs with P-256 at N = 10^7
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 2 --e2e-steps 2 --no-cpu > gpurun_out/s6_bench_1m.log 2> gpurun_out/s6_bench_1m.err; echo "bench rc=$?"
tail -c 300 gpurun_out/s6_bench_1m.err
