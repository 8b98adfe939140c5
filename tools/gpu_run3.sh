#!/bin/bash
# round-2 GPU session 3: full GPU parity suite, headline bench (N = 10^6) with host timeline, ncu of k_exp_fixed at N = 10^6
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python bench.py --n 4000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu --no-other > gpurun_out/s3_bench_smoke.log 2> gpurun_out/s3_bench_smoke.err; echo "smoke rc=$?"
timeout 300 python tools/prof_driver.py 1000000 3072 > gpurun_out/s3_prof_1m.log 2>&1
timeout 1500 python bench.py --steps 3 --warmup 2 --trace gpurun_out/s3_trace_1m.json > gpurun_out/s3_bench_1m.log 2> gpurun_out/s3_bench_1m.err; echo "bench rc=$?"
timeout 1800 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/s3_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s3_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_exp_fixed' -s 1 -c 1 -o gpurun_out/s3_prof_fixed_1m python tools/ncu_driver_fixed.py 1000000 > gpurun_out/s3_ncu1.log 2>&1; echo "ncu1 rc=$?"
