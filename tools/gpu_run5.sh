#!/bin/bash
# round-2 GPU session 5 (1 GPU): full GPU parity suite, then the headline bench as the driver runs it (shorter)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/s5_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s5_pytest_gpu.log
timeout 1500 python bench.py --steps 6 --warmup 3 > gpurun_out/s5_bench_1m.log 2> gpurun_out/s5_bench_1m.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s5_bench_ref.log 2> gpurun_out/s5_bench_ref.err; echo "ref rc=$?"
