#!/bin/bash
# 8-GPU headline only (strong scaling of N = 10^6), minimal steps: the budget left allows ~2 minutes on 8 GPUs
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 115 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 2 --warmup 2 --e2e-steps 2 --no-cpu --no-other > gpurun_out/m8_bench.log 2> gpurun_out/m8_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/m8_bench.err
