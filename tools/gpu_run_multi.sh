#!/bin/bash
# multi-GPU session ($1 = number of GPUs): NCCL parity tests (2 GPUs only) and the headline bench, strong scaling
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m${N}_smi.txt
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests/test_gpu_parallel.py -x -q > gpurun_out/m${N}_pytest_par.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m${N}_pytest_par.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu --trace gpurun_out/m${N}_trace.json > gpurun_out/m${N}_bench.log 2> gpurun_out/m${N}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/m${N}_bench.err
