#!/bin/bash
# round-2 GPU session 4 (1 GPU): headline bench with per-step times and the host timeline of a device-resident step
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "random_sources or group_ops or edge_cases or ring_ops" > gpurun_out/s4_pytest_codec.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s4_pytest_codec.log
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu --no-other --trace gpurun_out/s4_trace_1m.json > gpurun_out/s4_bench_1m.log 2> gpurun_out/s4_bench_1m.err; echo "bench rc=$?"
VMX_TRACE=1 timeout 600 python tools/trace_device_step.py 1000000 gpurun_out/s4_trace_device.json > gpurun_out/s4_trace_device.log 2>&1; echo "trace rc=$?"
