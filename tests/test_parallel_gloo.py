"""Multi-process (world_size 2 and 3, gloo) tests of the sharded array layer on the host-emulation
build: every sharded operation, a whole PoSBasicTW prove/verify and the decryption proof are
bit-identical to the single-process run (tests/parallel_worker.py does the asserting)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world: int, bits, n: int, emul_lib: str):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), VMX_LIBRARY_PATH=emul_lib, OMP_NUM_THREADS="1",
                   VMX_BUFFER_MIN="512")   # published messages take the pooled-buffer (memoryview) path as at full size
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "parallel_worker.py"), "cpu",
                                       str(bits), str(n)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out[-4000:])
    assert "PARALLEL OK" in outs[0]


@pytest.mark.parametrize("world,n", [(2, 24), (3, 20), (2, 1)])   # ragged and equal shards
def test_sharded_equals_single(emul_lib, world, n):
    _run(world, 512, n, emul_lib)


@pytest.mark.parametrize("world,n", [(2, 21)])
def test_sharded_curve_equals_single(emul_lib, world, n):
    _run(world, "P-256", n, emul_lib)
