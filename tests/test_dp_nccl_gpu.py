"""NCCL data-parallel parity on real GPUs: needs >= 2 devices (skipped on a 1-GPU box).
Runs tools/dp_check.py under torchrun: sharded training with all-reduced gradients must match a
single-GPU replica fed the whole batch."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_training_matches_full_batch(gpu):
    n = gpu.lib().tcr_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"], res
