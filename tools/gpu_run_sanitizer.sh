#!/bin/bash
# compute-sanitizer over a small selection of the GPU parity suite (one tool per gpurun call: $1 = memcheck | racecheck)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
tool=${1:-memcheck}
sel="test_edge_cases and 512 or test_group_ops and 512-300 or test_ring_ops and 512 or test_random_sources and 512 or test_transcript_parity and 512-1 or test_decryption_parity and 512 or test_ec_group_ops and P-256-40 or test_ec_transcript_parity and 1 or test_dedicated_squaring and 2048 or test_one_context_two_threads and 512"
# plain run first (the sanitizer only runs on a selection that passes without it)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "$sel" > gpurun_out/san_${tool}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/san_${tool}_plain.log; exit 1; }
tail -2 gpurun_out/san_${tool}_plain.log
timeout 2400 compute-sanitizer --tool $tool --log-file gpurun_out/san_${tool}.log --print-limit 50 python -m pytest tests/test_gpu_parity.py -x -q -k "$sel" > gpurun_out/san_${tool}_pytest.log 2>&1; echo "sanitizer rc=$?"
tail -3 gpurun_out/san_${tool}_pytest.log
tail -15 gpurun_out/san_${tool}.log
