"""Fused SyncBatchNorm (row N2): structure checks on one GPU, and -- when the box has at least two GPUs -- the
multi-process comparison against torch.nn.SyncBatchNorm (tests/dist_syncbn_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_convert_is_identity_without_process_group():
    from cloud_transformers_b200 import syncbn
    m = nn.Sequential(nn.Conv1d(3, 8, 1), nn.BatchNorm1d(8)).cuda()
    assert syncbn.convert_sync_batchnorm(m) is m and isinstance(m[1], nn.BatchNorm1d)


def test_fused_syncbn_matches_nccl_syncbn_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, TORCH_NCCL_ASYNC_ERROR_HANDLING="0")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577",
                          os.path.join(ROOT, "tests", "dist_syncbn_check.py")], capture_output=True, text=True, env=env,
                         timeout=600)
    assert out.returncode == 0 and "syncbn check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
