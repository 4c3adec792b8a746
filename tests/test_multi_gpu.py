"""Data-parallel parity on real GPUs (needs >= 2 devices; the 1-GPU round-end run skips it).
Run by hand with: gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("prec,exchange", [("fp32", "nccl"), ("bf16", "nccl"), ("bf16", "sharded"), ("bf16", "stale"), ("bf16", "state"), ("bf16", "p2p"), ("bf16", "chain"), ("bf16", "nvls"), ("bf16", "split")])
def test_two_rank_pcd_matches_single_gpu(prec, exchange):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py"), prec, exchange]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if "MGPU-SKIP" in out.stdout:   # NVSwitch multicast not available on this box: the worker says why
        pytest.skip([line for line in out.stdout.splitlines() if line.startswith("MGPU-SKIP")][0])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MGPU" in out.stdout
    print("\n".join(line for line in out.stdout.splitlines() if line.startswith("MGPU")))   # evidence for profiles/ (run with -s)
