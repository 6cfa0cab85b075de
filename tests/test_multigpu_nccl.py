"""N > 1 on real GPUs: one process per GPU over NCCL (tests/_nccl_worker.py). Skipped with fewer than 2 devices."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_ranks(world, p2p):
    """p2p = 1: ghost messages through peer memory (CUDA IPC, pack kernel storing into the neighbour's buffer, arrival flags);
    p2p = 0: NCCL send/recv. Same results either way: exchange bit-exact, trajectories to 1e-12 against the oracle."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=dict(os.environ, SPB_P2P=p2p))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" ok") == world
    assert out.stdout.count("p2p=" + p2p) == world
