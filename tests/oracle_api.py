"""ctypes view of the CPU oracles (oracle/jet_oracle.h).  TEST INFRASTRUCTURE.

Two interchangeable libraries export the same `jeto_*` functions:
  * oracle/libjetport.so      — plain-C restatement (kind "port"), built by `make -C oracle port`
  * oracle/_ref/libjetref.so  — the unmodified Jet reference behind a shim (kind "reference"),
                                built by `make -C oracle ref` where /root/reference exists.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_PATH = os.path.join(ROOT, "oracle", "libjetport.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libjetref.so")

from jetcu_loader import load_package

_pkg = load_package()
from fluid_engine_dev_b200.jetcu import (APIC, FLIP, PIC, PD, PF, PU64, ROW_DTYPE, ColliderShape, Grid, HybridApi, HybridParams,  # noqa: E402,F401
                                          LinearApi, MgpcgParams, MgpcgResult, Row, SmokeApi, SmokeParams, StepStats, _p, c64,
                                          f64, level_dims, size3, u32, u64)


class Oracle(LinearApi):
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        self.lib.jeto_kind.restype = C.c_char_p
        self.lib.jeto_mg_num_levels.restype = u32
        self.lib.jeto_mg_num_levels.argtypes = [u64, u64, u64, u32]
        self.lib.jeto_hybrid_num_particles.restype = u64
        self.lib.jeto_hybrid_num_particles.argtypes = [C.c_void_p]
        self.kind = self.lib.jeto_kind().decode()

    def _fn(self, name, raw=False):
        return getattr(self.lib, "jeto_" + name)

    def _call(self, name, *args, raw=False):
        fn = getattr(self.lib, "jeto_" + name)
        if raw:
            return fn(*args)
        fn.restype = C.c_int
        rc = fn(*args)
        if rc != 0:
            raise RuntimeError("jeto_%s failed rc=%d" % (name, rc))
        return rc

    def set_num_threads(self, n):
        self.lib.jeto_set_num_threads(int(n))

    def num_threads(self):
        return int(self.lib.jeto_num_threads())

    def bin_points_serial(self, positions, res, spacing):
        n = positions.shape[0]
        bucket = np.zeros(n, dtype=np.uint64)
        rank = np.zeros(n, dtype=np.uint64)
        self._call("bin_points_serial", u64(n), _p(c64(positions)), u64(res[0]), u64(res[1]), u64(res[2]),
                   f64(spacing), _p(bucket, PU64), _p(rank, PU64))
        return bucket, rank

    def hybrid(self, params):
        h = C.c_void_p()
        self._call("hybrid_create", C.byref(params), C.byref(h))
        return HybridApi(self._call, h, params)

    def smoke(self, params):
        h = C.c_void_p()
        self._call("smoke_create", C.byref(params), C.byref(h))
        return SmokeApi(self._call, h, params)


def load_port():
    return Oracle(PORT_PATH)


def load_ref():
    return Oracle(REF_PATH)


def have_ref():
    return os.path.exists(REF_PATH)


def have_port():
    return os.path.exists(PORT_PATH)
