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


@pytest.mark.parametrize("halo", ["nccl", "p2p"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_match_single_gpu(world, halo, tmp_path):
    """halo = "p2p": the same checks with JETCU_P2P_HALO=1 — halo planes pushed into the neighbours' CUDA-IPC mailboxes by
    kernels instead of NCCL send/recv; the result must not change by a bit and the path must really be the one in use."""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    n = 64
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29511 + world + (10 if halo == "p2p" else 0)), os.path.join(ROOT, "tests", "dist_worker.py"), str(n), "3"]
    env = dict(os.environ, JETCU_DIST_OUT=str(tmp_path), JETCU_P2P_HALO="1" if halo == "p2p" else "0")
    # own process group + a hard limit: a communication bug must fail this test, not hang the box
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, env=env, start_new_session=True)
    try:
        out, err = p.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        pytest.fail("distributed worker timed out at world size %d\n%s\n%s" % (world, out[-3000:], err[-3000:]))
    r = subprocess.CompletedProcess(cmd, p.returncode, out, err)
    files = [os.path.join(str(tmp_path), "rank%d.json" % q) for q in range(world)]
    assert r.returncode == 0 and all(os.path.exists(f) for f in files), (r.stdout[-4000:], r.stderr[-4000:])
    lines = [json.load(open(f)) for f in files]
    for l in lines:
        assert l["ok"], l
        assert l["comm"]["p2p"] == (halo == "p2p"), l["comm"]
