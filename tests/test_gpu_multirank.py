"""Multi-GPU parity (needs >= 2 GPUs; skipped on a one-GPU box): view sharding over NCCL ranks reproduces the single-GPU
images bit-for-bit, all_reduce_gradients equals the single-GPU sum, broadcast_scene replicates the scene
(tests/tools/multirank_check.py under torchrun).  Reference shape: DDP, one process per GPU (/root/reference/src/main.py:117-130)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_view_sharding_over_nccl_ranks_matches_one_gpu():
    n = min(torch.cuda.device_count(), 8)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "tools", "multirank_check.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"MULTIRANK OK {n}" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
