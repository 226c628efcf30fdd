"""Multi-rank parity over NCCL, run by pytest: spawns tests/multi_gpu_check.py under torchrun with two ranks when
at least two GPUs are visible (skipped otherwise; bench.py carries the same checks in its N > 1 line).
Checks: fine-tune gather + replicated / reduce-scatter backward, enqueue of the gathered keys (eager, deferred,
CUDA-graph replay), gallery-sharded eval vs one GPU, SyncBatchNorm MLP -- each against the numpy oracle on the
rank-major concatenation of the ranks' inputs (modules/modeling.py:25-36, 244-284, 698-709)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_multi_rank_checks_over_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and ("multi_gpu_check W=%d ok" % world) in p.stdout, p.stdout[-6000:]
