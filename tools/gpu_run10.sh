#!/bin/bash
# round-2 GPU session 10 (1 GPU, the last 2.7 GPU-minutes of the round): the session-type / pre-computation tests on the CUDA build
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 130 python -m pytest tests/test_gpu_sessions.py -x -q --durations=5 > gpurun_out/s10_pytest_sessions.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/s10_pytest_sessions.log
