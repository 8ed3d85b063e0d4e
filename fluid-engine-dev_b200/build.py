"""Builds libjetcu.so (the C-ABI product library) in-tree with nvcc for sm_100a only.

    python fluid-engine-dev_b200/build.py [--force]

No torch, no JIT cache: the .so lands next to this file so it travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjetcu.so")
SOURCES = ["abi.cu", "fdm.cu", "iccg.cu", "pressure.cu", "particles.cu", "hybrid.cu", "smoke.cu", "diffusion.cu", "comm.cu", "io.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                                  # replay the reference's non-contracted IEEE arithmetic
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v" if os.environ.get("JETCU_PTXAS_V") else "-O3",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


STAMP = LIB + ".srchash"


def source_hash():
    """sha256 over the compiler flags and the CONTENT of every file the library is built from (not their mtimes)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "jetcu.h")]
    for d in deps:
        if os.path.isfile(d):
            h.update(os.path.basename(d).encode())
            h.update(open(d, "rb").read())
    return h.hexdigest()


def needs_build():
    """The in-tree library is reused only if it was built from exactly these sources (content hash recorded next to it)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    odir = os.path.join(HERE, "build")
    os.makedirs(odir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    for src in sources():
        obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc_path()] + flags + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for " + cmd[-3])
    libs = ["-lcudart", "-ldl"]   # NCCL is resolved with dlopen at run time (csrc/comm.cu): no link-time dependency
    cmd = [nvcc_path(), "-shared", "-o", LIB] + objs + libs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
