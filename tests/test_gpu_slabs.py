"""Multi-GPU parity: the z-slab distributed path must reproduce the single-GPU result bit for bit.
Needs >= 2 GPUs on the box; launched through torchrun exactly as bench.py is (one rank per GPU, NCCL)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_match_single_gpu(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    n = 64
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29511 + world), os.path.join(ROOT, "tests", "dist_worker.py"), str(n), "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == world, (r.stdout[-4000:], r.stderr[-4000:])
    for l in lines:
        assert l["ok"], l
