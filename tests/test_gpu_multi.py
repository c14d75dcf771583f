"""GPU (>= 2 devices): frame sharding + NCCL all-reduce of the grids equals the sequential result."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_frame_sharded_replay_over_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for r in range(world):
        assert "rank %d/%d ok" % (r, world) in out.stdout
