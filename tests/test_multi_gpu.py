"""Sharded ARS iteration on real GPUs (skipped on boxes with one GPU): launches tests/dist_check.py under
torchrun with 2 ranks (and with every GPU of the box when there are more), under a watchdog."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
    assert out.returncode == 0, out.stdout[-4000:]
    assert out.stdout.count("dist_check ok") == 9, out.stdout[-4000:]   # 2 transports x 4 cases + the shard="auto" check


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_iteration_matches_single_gpu_2_ranks():
    _run(2, 29731)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs >= 4 GPUs")
def test_sharded_iteration_matches_single_gpu_all_ranks():
    n = torch.cuda.device_count()
    _run(8 if n >= 8 else 4, 29741)
