"""Builds tests/hostcheck/libhostcheck.so: the product's __host__ __device__ kernel bodies compiled for the CPU
(nvcc, host code only).  TEST INFRASTRUCTURE."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhostcheck.so")
SRC = os.path.join(HERE, "hostcheck.cu")
CSRC = os.path.join(HERE, "..", "..", "fluid-engine-dev_b200", "csrc")


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    return None


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("collider.cuh", "collider_ops.cuh", "particle_ops.cuh", "p2g_ops.cuh", "smoke_ops.cuh", "fdm_ops.cuh", "pressure_ops.cuh", "common.cuh", "sampling.cuh", "diffusion_ops.cuh", "iccg_ops.cuh", "p2p_ops.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        return None
    # host arithmetic must not be contracted into FMAs (the device side is built with -fmad=false)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false", "-Xcompiler",
           "-fPIC,-ffp-contract=off", "-shared", "-o", LIB, SRC]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostcheck build failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build())
