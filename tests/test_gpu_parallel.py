"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the NCCL-sharded run of every
array operation, PoSBasicTW, the decryption proof and the Fiat-Shamir shuffle session is
bit-identical to the single-GPU run (tests/parallel_worker.py asserts)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("bits,n", [(512, 3000), (3072, 20000), ("P-256", 50000)])
def test_nccl_sharded_equals_single(bits, n):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4
    env = dict(os.environ)
    env.pop("VMX_LIBRARY_PATH", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "parallel_worker.py"), "gpu", str(bits), str(n)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout + r.stderr)[-6000:]
    assert "PARALLEL OK" in r.stdout
