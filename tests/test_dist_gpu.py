"""Multi-GPU (needs >= 2 GPUs on the box: `gpurun --gpus 2`): data-parallel parity of the reference's train step
(tools/dp_parity.py) — sharded replicas with all-reduced gradient buckets, agreed branch flag and SyncBatchNorm give
the gradients of the single-process full-batch step."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("domain,n_lo,ratio", [("fluid", 256, 4), ("action", 128, 16)])
def test_data_parallel_gradients_equal_single_process(domain, n_lo, ratio):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "dp_parity.py"), "--domain", domain, "--batch", "4",
           "--n-lo", str(n_lo), "--ratio", str(ratio)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stderr[-2000:]
    out = json.loads(lines[-1])
    assert r.returncode == 0 and out["ok"], out
