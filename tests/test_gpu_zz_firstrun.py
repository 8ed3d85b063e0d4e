"""First-GPU-run files (tests/firstrun/zz_*.py), each executed in a CHILD pytest process.

Those files exercise CUDA paths written after round 1's GPU minutes were spent (analytic colliders, the smoke solver, the
full-size property checks).  Their kernel bodies are verified bit-for-bit on the CPU (tests/test_hostcheck_*.py) but the
launches themselves have never run, so they are kept at arm's length from the established parity suite:
  * a child process — an illegal address, a sticky CUDA error or a segfault ends the child, not this run;
  * a hard timeout on the child;
  * non-strict xfail here — a clean child shows up as XPASS, anything else as XFAIL with the log's tail as the reason;
  * this file sorts last.
The child's full log goes to gpurun_out/firstrun/<name>.log.  Round 2 moves the files back under tests/ once green."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = {"collider": 900, "smoke": 900, "fuzz": 900, "fullsize": 1500}           # name -> child time limit, seconds

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first GPU run (see module docstring)")]


@pytest.mark.timeout(1600, method="thread")
@pytest.mark.parametrize("name", list(FILES))
def test_firstrun_file_in_child(name):
    path = os.path.join(ROOT, "tests", "firstrun", "zz_%s.py" % name)
    logdir = os.path.join(ROOT, "gpurun_out", "firstrun")
    os.makedirs(logdir, exist_ok=True)
    log = os.path.join(logdir, name + ".log")
    cmd = [sys.executable, "-m", "pytest", path, "-q", "-m", "gpu", "-rA", "--tb=short", "-p", "no:cacheprovider"]
    with open(log, "w") as f:
        child = subprocess.Popen(cmd, cwd=ROOT, stdout=f, stderr=subprocess.STDOUT, start_new_session=True)
        try:
            rc = child.wait(timeout=FILES[name])
        except subprocess.TimeoutExpired:
            os.killpg(child.pid, 9)                                  # the exact process group started above
            child.wait()
            rc = "timeout after %d s" % FILES[name]
    tail = "".join(open(log, errors="replace").readlines()[-25:])
    assert rc == 0, "child pytest on %s: rc=%s\n%s" % (os.path.relpath(path, ROOT), rc, tail)
