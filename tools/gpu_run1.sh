#!/bin/bash
# round-2 GPU session 1: bench smoke, GPU parity suite, headline bench at N = 10^6, launch list, per-kernel timings
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/s1_smi.txt
nproc >> gpurun_out/s1_smi.txt
timeout 300 python bench.py --n 4000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu --no-other > gpurun_out/s1_bench_smoke.log 2> gpurun_out/s1_bench_smoke.err; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/s1_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s1_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 --phases > gpurun_out/s1_bench_1m.log 2> gpurun_out/s1_bench_1m.err; echo "bench rc=$?"
timeout 300 python tools/prof_driver.py 100000 > gpurun_out/s1_prof_100k.log 2>&1
timeout 300 python tools/prof_driver.py 1000000 > gpurun_out/s1_prof_1m.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s1_launches_100k.csv python bench.py --n 100000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-other > gpurun_out/s1_ncu_launches.log 2>&1; echo "ncu rc=$?"
