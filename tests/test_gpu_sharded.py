"""GPU, 2 ranks over NCCL: the row-sharded trainer reproduces the single-GPU module path (needs >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind,peer", [("custom", "1"), ("pyg", "1"), ("pyg", "0"), ("custom", "auto")])
def test_sharded_matches_single_gpu(kind, peer):
    """peer = "1": row exchanges pull from peer memory (b200gat_peer_*); "0": NCCL all-gathers."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "sharded_check.py"), kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, B200GAT_PEER=peer))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_OK" in r.stdout
    # a box without CUDA IPC between the two processes falls back to NCCL and says so on stderr
    assert ("exchange=nccl" if peer == "0" else "exchange=peer") in r.stdout or "peer exchange unavailable" in r.stderr, \
        r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_world1_matches_modules():
    """World size 1 (no collectives): the manual forward/backward equals the autograd module path bit for bit."""
    assert torch.cuda.is_available()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sharded_check.py"), "pyg"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_OK" in r.stdout
