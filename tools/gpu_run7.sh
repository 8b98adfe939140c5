#!/bin/bash
# round-2 GPU session 7 (1 GPU, ~2 minutes): the native verifier on the CUDA build + smoke()
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "native_vmnv" > gpurun_out/s7_pytest_native.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s7_pytest_native.log
timeout 120 python bench.py --workload verify-mix --n 100000 --steps 2 --warmup 1 --no-cpu > gpurun_out/s7_bench_mix.log 2> gpurun_out/s7_bench_mix.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s7_bench_mix.err
