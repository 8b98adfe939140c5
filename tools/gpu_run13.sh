#!/bin/bash
# round-2 GPU session 13 (1 GPU, the last GPU seconds of the round): sanity of the rebuilt libvmx.so (host-side argument
# checks only) and of the final host code on the session tests
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 36 python -m pytest tests/test_gpu_sessions.py -x -q > gpurun_out/s13_pytest_sessions.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s13_pytest_sessions.log
