"""ctypes binding of the jetcu C ABI (include/jetcu.h) — the Python host side of the B200-native
hybrid-liquid path.  Arrays are numpy, Jet Array3 layout: shape (nz, ny, nx), x fastest.

There is no CPU fallback: `Jetcu()` raises if libjetcu.so is missing or no sm_100 device is usable.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjetcu.so")

u64 = C.c_uint64
u32 = C.c_uint32
f64 = C.c_double
PD = C.POINTER(C.c_double)
PF = C.POINTER(C.c_float)
PU64 = C.POINTER(C.c_uint64)


class Row(C.Structure):
    _fields_ = [("center", f64), ("right", f64), ("up", f64), ("front", f64)]


ROW_DTYPE = np.dtype([("center", "<f8"), ("right", "<f8"), ("up", "<f8"), ("front", "<f8")])


class MgpcgParams(C.Structure):
    _fields_ = [("max_iterations", u32), ("max_levels", u32), ("n_restriction", u32), ("n_correction", u32),
                ("n_coarsest", u32), ("n_final", u32), ("tolerance", f64), ("sor_factor", f64),
                ("use_red_black", C.c_int32), ("linear_solver", C.c_int32)]

    @classmethod
    def make(cls, max_iterations=100, max_levels=5, n_restriction=5, n_correction=5, n_coarsest=20, n_final=20,
             tolerance=1e-9, sor_factor=1.5, use_red_black=True):
        return cls(max_iterations, max_levels, n_restriction, n_correction, n_coarsest, n_final, tolerance,
                   sor_factor, 1 if use_red_black else 0, 0)

    @classmethod
    def iccg(cls, max_iterations=100, tolerance=1e-6, compressed=False):
        """FdmIccgSolver3(maxNumberOfIterations, tolerance) — Jet's default linear solver — as the pressure solve's solver;
        compressed = GridFluidSolver3::setUseCompressedLinearSystem(true) (the device solves the same system uncompressed)."""
        return cls(max_iterations, 1, 0, 0, 0, 0, tolerance, 1.0, 1, 1 | (2 if compressed else 0))


class MgpcgResult(C.Structure):
    _fields_ = [("iterations", u32), ("returned", C.c_int32), ("last_residual", f64), ("true_residual_l2", f64),
                ("b_l2", f64), ("solve_ms", f64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Grid(C.Structure):
    _fields_ = [("nx", u64), ("ny", u64), ("nz", u64), ("hx", f64), ("hy", f64), ("hz", f64),
                ("ox", f64), ("oy", f64), ("oz", f64)]

    @classmethod
    def cube(cls, n, domain=1.0, origin=(0.0, 0.0, 0.0)):
        h = domain / n
        return cls(n, n, n, h, h, h, *origin)

    @property
    def res(self):
        return (int(self.nx), int(self.ny), int(self.nz))

    @property
    def usize(self):
        return (int(self.nx) + 1, int(self.ny), int(self.nz))

    @property
    def vsize(self):
        return (int(self.nx), int(self.ny) + 1, int(self.nz))

    @property
    def wsize(self):
        return (int(self.nx), int(self.ny), int(self.nz) + 1)


class ColliderShape(C.Structure):
    """jetcu_collider_shape / jeto_collider_shape (include/jetcu.h)."""
    _fields_ = [("kind", C.c_int32), ("flipped", C.c_int32), ("a", f64 * 3), ("b", f64 * 3)]


class HybridParams(C.Structure):
    _fields_ = [("grid", Grid), ("mode", C.c_int32), ("closed_domain_boundary_flag", C.c_int32),
                ("pic_blending_factor", f64), ("gravity", f64 * 3), ("max_cfl", f64), ("mgpcg", MgpcgParams)]

    @classmethod
    def make(cls, grid, mode=1, mgpcg=None, pic_blending_factor=0.0, gravity=(0.0, -9.8, 0.0), max_cfl=5.0,
             closed_domain_boundary_flag=63):
        p = cls()
        p.grid = grid
        p.mode = mode
        p.closed_domain_boundary_flag = closed_domain_boundary_flag
        p.pic_blending_factor = pic_blending_factor
        p.gravity = (f64 * 3)(*gravity)
        p.max_cfl = max_cfl
        p.mgpcg = mgpcg if mgpcg is not None else MgpcgParams.make()
        return p


class StepStats(C.Structure):
    _fields_ = [("pcg_iterations", u32), ("pcg_last_residual", f64), ("pcg_true_residual_l2", f64),
                ("pcg_b_l2", f64), ("ms_total", f64), ("ms_bin", f64), ("ms_p2g", f64), ("ms_sdf", f64),
                ("ms_extrapolate", f64), ("ms_bc", f64), ("ms_pressure", f64), ("ms_g2p", f64), ("ms_move", f64),
                ("ms_mgpcg_solve", f64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


PIC, FLIP, APIC = 0, 1, 2


class SmokeParams(C.Structure):
    """== the knobs of jet::GridSmokeSolver3 (include/jet/grid_smoke_solver3.h:170-175) + GridFluidSolver3's."""
    _fields_ = [("grid", Grid), ("closed_domain_boundary_flag", C.c_int32), ("advection", C.c_int32), ("gravity", f64 * 3),
                ("max_cfl", f64), ("buoyancy_density_factor", f64), ("buoyancy_temperature_factor", f64),
                ("density_decay", f64), ("temperature_decay", f64), ("mgpcg", MgpcgParams)]

    @classmethod
    def make(cls, grid, mgpcg=None, gravity=(0.0, -9.8, 0.0), max_cfl=5.0, closed_domain_boundary_flag=63,
             buoyancy_density_factor=-0.000625, buoyancy_temperature_factor=5.0, density_decay=0.001, temperature_decay=0.001,
             advection="cubic"):
        p = cls()
        p.grid = grid
        p.advection = {"cubic": 0, "linear": 1}[advection]   # CubicSemiLagrangian3 (Jet's default) / SemiLagrangian3
        p.closed_domain_boundary_flag = closed_domain_boundary_flag
        p.gravity = (f64 * 3)(*gravity)
        p.max_cfl = max_cfl
        p.buoyancy_density_factor = buoyancy_density_factor
        p.buoyancy_temperature_factor = buoyancy_temperature_factor
        p.density_decay = density_decay
        p.temperature_decay = temperature_decay
        p.mgpcg = mgpcg if mgpcg is not None else MgpcgParams.make()
        return p


class SmokeApi:
    """Wrapper over a jeto_smoke* / jetcu_smoke* handle (GridSmokeSolver3); `lib_call(name, *args)` raises on error."""

    def __init__(self, lib_call, handle, params):
        self._c = lib_call
        self.h = handle
        self.params = params
        self.grid = params.grid

    def close(self):
        if self.h is not None:
            self._c("smoke_destroy", self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def emit_step(self, shapes, density=(0.0, 1.0), temperature=(0.0, 1.0), as_set=None):
        """VolumeGridEmitter3 with step-function targets on smoke density and temperature (the source of
        src/examples/smoke_sim), fired once: value = max(old, (max - min) * (1 - smearedHeaviside(sdf / min h)) + min).
        `shapes`: dicts as for HybridApi.set_collider_shapes."""
        if as_set is None:
            as_set = len(shapes) != 1
        arr = (ColliderShape * max(len(shapes), 1))()
        for i, sh in enumerate(shapes):
            arr[i].kind = HybridApi.SHAPE_KINDS[sh["kind"]]
            arr[i].flipped = 1 if sh.get("flipped", False) else 0
            arr[i].a = (f64 * 3)(*sh["a"])
            b = tuple(sh.get("b", (0.0, 0.0, 0.0))) + (0.0, 0.0, 0.0)
            arr[i].b = (f64 * 3)(*b[:3])
        self._c("smoke_emit_step", self.h, arr, u32(len(shapes)), C.c_int(1 if as_set else 0), f64(density[0]), f64(density[1]),
                f64(temperature[0]), f64(temperature[1]))

    def set_collider_shapes(self, shapes, as_set=None, linear_velocity=(0.0, 0.0, 0.0), angular_velocity=(0.0, 0.0, 0.0), friction=0.0):
        """GridFluidSolver3::setCollider(RigidBodyCollider3 over analytic surfaces); shapes as for HybridApi.set_collider_shapes,
        an empty list removes the collider."""
        if as_set is None:
            as_set = len(shapes) > 1
        arr = (ColliderShape * max(len(shapes), 1))()
        for i, sh in enumerate(shapes):
            arr[i].kind = HybridApi.SHAPE_KINDS[sh["kind"]]
            arr[i].flipped = 1 if sh.get("flipped", False) else 0
            arr[i].a = (f64 * 3)(*sh["a"])
            b = tuple(sh.get("b", (0.0, 0.0, 0.0))) + (0.0, 0.0, 0.0)
            arr[i].b = (f64 * 3)(*b[:3])
        self._c("smoke_set_collider_shapes", self.h, arr, u32(len(shapes)), C.c_int(1 if as_set else 0), (f64 * 3)(*linear_velocity),
                (f64 * 3)(*angular_velocity), f64(friction))

    def set_fields(self, density=None, temperature=None, u=None, v=None, w=None):
        self._c("smoke_set_fields", self.h, _p(c64(density)), _p(c64(temperature)), _p(c64(u)), _p(c64(v)), _p(c64(w)))

    def get_fields(self):
        g = self.grid
        out = dict(density=np.zeros(size3(g.res)), temperature=np.zeros(size3(g.res)), u=np.zeros(size3(g.usize)),
                   v=np.zeros(size3(g.vsize)), w=np.zeros(size3(g.wsize)), pressure=np.zeros(size3(g.res)))
        self._c("smoke_get_fields", self.h, _p(out["density"]), _p(out["temperature"]), _p(out["u"]), _p(out["v"]), _p(out["w"]),
                _p(out["pressure"]))
        return out

    def step(self, dt):
        r = MgpcgResult()
        self._c("smoke_step", self.h, f64(dt), C.byref(r))
        return r

    def last_step_ms(self):
        f, p, a, d = f64(0), f64(0), f64(0), f64(0)
        self._c("smoke_last_step_ms", self.h, C.byref(f), C.byref(p), C.byref(a), C.byref(d))
        return {"ms_forces": f.value, "ms_pressure": p.value, "ms_advection": a.value, "ms_decay": d.value}

    def cfl(self, dt):
        c = f64(0.0)
        self._c("smoke_cfl", self.h, f64(dt), C.byref(c))
        return float(c.value)

    def num_sub_steps(self, dt):
        n = u32(0)
        self._c("smoke_num_sub_steps", self.h, f64(dt), C.byref(n))
        return int(n.value)

    def compute_external_forces(self, dt):
        self._c("smoke_compute_external_forces", self.h, f64(dt))

    def compute_pressure(self, dt):
        r = MgpcgResult()
        self._c("smoke_compute_pressure", self.h, f64(dt), C.byref(r))
        return r

    def compute_advection(self, dt):
        self._c("smoke_compute_advection", self.h, f64(dt))

    def end_advance_time_step(self, dt):
        self._c("smoke_end_advance_time_step", self.h, f64(dt))

    def set_diffusion(self, viscosity=0.0, smoke_diffusion=0.0, temperature_diffusion=0.0):
        """GridFluidSolver3::setViscosityCoefficient, GridSmokeSolver3::setSmokeDiffusionCoefficient /
        setTemperatureDiffusionCoefficient (> kEpsilonD switches the GridBackwardEulerDiffusionSolver3 solves on)."""
        self._c("smoke_set_diffusion", self.h, f64(viscosity), f64(smoke_diffusion), f64(temperature_diffusion))

    def compute_viscosity(self, dt):
        self._c("smoke_compute_viscosity", self.h, f64(dt))


def size3(shape_xyz):
    """numpy arrays holding Jet Array3 data are indexed [k, j, i] (x fastest)."""
    nx, ny, nz = shape_xyz
    return (nz, ny, nx)


def _p(a, typ=PD):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ)


def c64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def level_dims(nx, ny, nz, levels):
    return [(nx >> l, ny >> l, nz >> l) for l in range(levels)]


class LinearApi:
    """Methods shared by the oracle wrapper and the jetcu wrapper (same argument order).
    Subclasses provide self._call(name, *args) which prepends the ctx where needed."""

    def mg_num_levels(self, nx, ny, nz, max_levels):
        return int(self._fn("mg_num_levels", raw=True)(nx, ny, nz, max_levels))

    def mvm(self, A, v):
        nz, ny, nx = v.shape
        out = np.empty_like(v)
        self._call("fdm_mvm", u64(nx), u64(ny), u64(nz), _p(A, C.POINTER(Row)), _p(v), _p(out))
        return out

    def residual(self, A, x, b):
        nz, ny, nx = x.shape
        out = np.empty_like(x)
        self._call("fdm_residual", u64(nx), u64(ny), u64(nz), _p(A, C.POINTER(Row)), _p(x), _p(b), _p(out))
        return out

    def relax_red_black(self, A, b, x, sor=1.5, iterations=1):
        nz, ny, nx = x.shape
        out = np.array(x, dtype=np.float64, order="C", copy=True)
        self._call("fdm_relax_red_black", u64(nx), u64(ny), u64(nz), _p(A, C.POINTER(Row)), _p(b), f64(sor),
                   u32(iterations), _p(out))
        return out

    def relax(self, A, b, x, sor=1.5, iterations=1):
        """FdmGaussSeidelSolver3::relax, the natural (lexicographic) ordering (fdm_gauss_seidel_solver3.cpp:103-120)."""
        nz, ny, nx = x.shape
        out = np.array(x, dtype=np.float64, order="C", copy=True)
        self._call("fdm_relax", u64(nx), u64(ny), u64(nz), _p(A, C.POINTER(Row)), _p(b), f64(sor), u32(iterations), _p(out))
        return out

    def restrict(self, finer):
        nz, ny, nx = finer.shape
        out = np.empty((nz // 2, ny // 2, nx // 2), dtype=np.float64)
        self._call("fdm_restrict", u64(nx), u64(ny), u64(nz), _p(finer), _p(out))
        return out

    def correct(self, coarser, finer):
        nz, ny, nx = finer.shape
        out = np.array(finer, dtype=np.float64, order="C", copy=True)
        self._call("fdm_correct", u64(nx), u64(ny), u64(nz), _p(coarser), _p(out))
        return out

    def dot(self, a, b):
        r = f64(0)
        self._call("fdm_dot", u64(a.size), _p(a), _p(b), C.byref(r))
        return r.value

    def axpy(self, a, x, y):
        out = np.empty_like(x)
        self._call("fdm_axpy", u64(x.size), f64(a), _p(x), _p(y), _p(out))
        return out

    @staticmethod
    def _level_ptrs(A_levels):
        arr = (C.POINTER(Row) * len(A_levels))()
        for i, a in enumerate(A_levels):
            assert a.flags["C_CONTIGUOUS"] and a.dtype == ROW_DTYPE
            arr[i] = a.ctypes.data_as(C.POINTER(Row))
        return arr

    def vcycle(self, A_levels, b, x, params):
        nz, ny, nx = b.shape
        out = np.array(x, dtype=np.float64, order="C", copy=True)
        self._call("fdm_vcycle", u64(nx), u64(ny), u64(nz), u32(len(A_levels)), self._level_ptrs(A_levels), _p(b),
                   _p(out), C.byref(params))
        return out

    def mgpcg_solve(self, A_levels, b, params):
        nz, ny, nx = b.shape
        x = np.zeros_like(b)
        res = MgpcgResult()
        self._call("mgpcg_solve", u64(nx), u64(ny), u64(nz), u32(len(A_levels)), self._level_ptrs(A_levels), _p(b),
                   _p(x), C.byref(params), C.byref(res))
        return x, res

    def iccg_solve(self, A, b, max_iterations=100, tolerance=1e-6):
        nz, ny, nx = b.shape
        x = np.zeros_like(b)
        res = MgpcgResult()
        A = np.ascontiguousarray(A)
        self._call("iccg_solve", u64(nx), u64(ny), u64(nz), A.ctypes.data_as(C.POINTER(Row)), _p(b), _p(x), u32(max_iterations),
                   f64(tolerance), C.byref(res))
        return x, res

    def iccg_precondition(self, A, b):
        nz, ny, nx = b.shape
        d, y, x = np.zeros_like(b), np.zeros_like(b), np.zeros_like(b)
        A = np.ascontiguousarray(A)
        self._call("iccg_precondition", u64(nx), u64(ny), u64(nz), A.ctypes.data_as(C.POINTER(Row)), _p(b), _p(d), _p(y), _p(x))
        return d, y, x

    def pressure_solve(self, grid, u, v, w, params, collider_sdf=None, bnd=(None, None, None), fluid_sdf=None,
                       out=None, want_pressure=True):
        if out is None:
            uo, vo, wo = u.copy(), v.copy(), w.copy()
        else:
            uo, vo, wo = (np.array(a, dtype=np.float64, order="C", copy=True) for a in out)
        p = np.zeros(size3(grid.res), dtype=np.float64) if want_pressure else None
        res = MgpcgResult()
        self._call("pressure_solve", C.byref(grid), _p(u), _p(v), _p(w), _p(c64(collider_sdf)), _p(c64(bnd[0])),
                   _p(c64(bnd[1])), _p(c64(bnd[2])), _p(c64(fluid_sdf)), C.byref(params), _p(uo), _p(vo), _p(wo),
                   _p(p), C.byref(res))
        return uo, vo, wo, p, res

    def diffusion_solve_faces(self, grid, u, v, w, coefficient, dt, boundary_type=1, collider_sdf=None, fluid_sdf=None,
                              max_iterations=100, tolerance=2.220446049250313e-16):
        """GridBackwardEulerDiffusionSolver3::solve(FaceCenteredGrid3) — boundary_type 0 Dirichlet, 1 Neumann (Jet's default);
        ICCG(100, kEpsilonD) unless told otherwise.  Returns u, v, w and the three solves' results."""
        uo, vo, wo = np.zeros_like(u), np.zeros_like(v), np.zeros_like(w)
        res = (MgpcgResult * 3)()
        self._call("diffusion_solve_faces", C.byref(grid), _p(c64(u)), _p(c64(v)), _p(c64(w)), f64(coefficient), f64(dt),
                   C.c_int(boundary_type), _p(c64(collider_sdf)), _p(c64(fluid_sdf)), u32(max_iterations), f64(tolerance),
                   _p(uo), _p(vo), _p(wo), res)
        return uo, vo, wo, list(res)

    def diffusion_solve_cells(self, grid, source, coefficient, dt, boundary_type=1, collider_sdf=None, fluid_sdf=None,
                              max_iterations=100, tolerance=2.220446049250313e-16, want_markers=False):
        """GridBackwardEulerDiffusionSolver3::solve(ScalarGrid3) on a cell-centred grid."""
        dest = np.zeros_like(source)
        markers = np.zeros(source.shape, dtype=np.int8) if want_markers else None
        res = MgpcgResult()
        self._call("diffusion_solve_cells", C.byref(grid), _p(c64(source)), f64(coefficient), f64(dt), C.c_int(boundary_type),
                   _p(c64(collider_sdf)), _p(c64(fluid_sdf)), u32(max_iterations), f64(tolerance), _p(dest),
                   markers.ctypes.data_as(C.c_void_p) if want_markers else None, C.byref(res))
        return (dest, res, markers) if want_markers else (dest, res)

    def pressure_get_system(self, grid, level):
        nx, ny, nz = (int(grid.nx) >> level, int(grid.ny) >> level, int(grid.nz) >> level)
        A = np.zeros((nz, ny, nx), dtype=ROW_DTYPE)
        b = np.zeros((nz, ny, nx), dtype=np.float64)
        sdf = np.zeros((nz, ny, nx), dtype=np.float32)
        uw = np.zeros((nz, ny, nx + 1), dtype=np.float32)
        vw = np.zeros((nz, ny + 1, nx), dtype=np.float32)
        ww = np.zeros((nz + 1, ny, nx), dtype=np.float32)
        self._call("pressure_get_system", u32(level), _p(A, C.POINTER(Row)), _p(b), _p(sdf, PF), _p(uw, PF),
                   _p(vw, PF), _p(ww, PF))
        return dict(A=A, b=b, fluid_sdf=sdf, u_weights=uw, v_weights=vw, w_weights=ww)

    def pressure_vcycle(self, b, x, params):
        """One V-cycle on the system built by the last pressure_solve (jetcu only; test hook)."""
        out = np.array(x, dtype=np.float64, order="C", copy=True)
        self._call("pressure_vcycle", _p(c64(b)), _p(out), C.byref(params))
        return out

    def bin_points(self, positions, res, spacing):
        n = positions.shape[0]
        nb = res[0] * res[1] * res[2]
        ku = np.zeros(n, dtype=np.uint64)
        si = np.zeros(n, dtype=np.uint64)
        ks = np.zeros(n, dtype=np.uint64)
        st = np.zeros(nb, dtype=np.uint64)
        en = np.zeros(nb, dtype=np.uint64)
        self._call("bin_points", u64(n), _p(c64(positions)), u64(res[0]), u64(res[1]), u64(res[2]), f64(spacing),
                   _p(ku, PU64), _p(si, PU64), _p(ks, PU64), _p(st, PU64), _p(en, PU64))
        return dict(keys_unsorted=ku, sorted_indices=si, keys_sorted=ks, start_table=st, end_table=en)


class HybridApi:
    """Wrapper over a jeto_hybrid* / jetcu_hybrid* handle; `lib_call(name, *args)` raises on error."""

    def __init__(self, lib_call, handle, params):
        self._c = lib_call
        self.h = handle
        self.params = params
        self.grid = params.grid

    def close(self):
        if self.h is not None:
            self._c("hybrid_destroy", self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_particles(self, pos, vel=None):
        pos = c64(pos)
        vel = c64(vel) if vel is not None else np.zeros_like(pos)
        self._c("hybrid_set_particles", self.h, u64(pos.shape[0]), _p(pos), _p(vel))

    def append_particles(self, pos, vel=None):
        pos = c64(pos)
        vel = c64(vel) if vel is not None else None
        self._c("hybrid_append_particles", self.h, u64(pos.shape[0]), _p(pos), _p(vel) if vel is not None else None)

    def emit_box(self, lower, upper, spacing, v0=(0.0, 0.0, 0.0)):
        n = u64(0)
        self._c("hybrid_emit_box", self.h, (f64 * 3)(*lower), (f64 * 3)(*upper), f64(spacing), (f64 * 3)(*v0),
                C.byref(n))
        return n.value

    def num_particles(self):
        return int(self._c("hybrid_num_particles", self.h, raw=True))

    def get_particles(self):
        n = self.num_particles()
        pos = np.zeros((n, 3))
        vel = np.zeros((n, 3))
        self._c("hybrid_get_particles", self.h, _p(pos), _p(vel))
        return pos, vel

    def save_positions(self, path, fmt="pos"):
        """saveParticleAsPos / saveParticleAsXyz of the reference's examples (src/examples/hybrid_liquid_sim/main.cpp:29-63),
        from the device-resident particle set; jetcu only (the oracles have no file output)."""
        self._c("hybrid_save_positions", self.h, C.c_char_p(os.fsencode(path)), C.c_int({"pos": 0, "xyz": 1}[fmt]))

    def num_local_particles(self):
        return int(self._c("hybrid_num_local_particles", self.h, raw=True))

    def get_local_particles(self):
        """This rank's own particles (slab mode; all of them on one GPU): ids, positions, velocities."""
        n = self.num_local_particles()
        ids = np.zeros(n, dtype=np.uint64)
        pos = np.zeros((n, 3))
        vel = np.zeros((n, 3))
        self._c("hybrid_get_local_particles", self.h, _p(ids, PU64), _p(pos), _p(vel))
        return ids, pos, vel

    def set_local_particles(self, ids, pos, vel, n_global):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        self._c("hybrid_set_local_particles", self.h, u64(ids.shape[0]), _p(ids, PU64), _p(c64(pos)), _p(c64(vel)),
                u64(n_global))

    def get_affine(self):
        n = self.num_particles()
        cx, cy, cz = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        self._c("hybrid_get_affine", self.h, _p(cx), _p(cy), _p(cz))
        return cx, cy, cz

    SHAPE_KINDS = {"box": 1, "sphere": 2, "cylinder": 3, "plane": 4}

    def _shape_array(self, shapes):
        arr = (ColliderShape * max(len(shapes), 1))()
        for i, sh in enumerate(shapes):
            arr[i].kind = self.SHAPE_KINDS[sh["kind"]]
            arr[i].flipped = 1 if sh.get("flipped", False) else 0
            arr[i].a = (f64 * 3)(*sh["a"])
            b = tuple(sh.get("b", (0.0, 0.0, 0.0))) + (0.0, 0.0, 0.0)
            arr[i].b = (f64 * 3)(*b[:3])
        return arr

    def emit_shapes(self, shapes, spacing, initial_velocity=(0.0, 0.0, 0.0), as_set=None, lattice="grid"):
        """VolumeParticleEmitter3 (one shot, jitter 0, GridPointGenerator3, maxRegion = the domain) over one analytic surface
        or an ImplicitSurfaceSet3 of them (dicts as for set_collider_shapes); returns the number of particles emitted."""
        if as_set is None:
            as_set = len(shapes) != 1
        n = u64(0)
        # lattice: "grid" = GridPointGenerator3 (what hybrid_liquid_sim installs), "bcc" = BccLatticePointGenerator (the default)
        self._c("hybrid_emit_shapes", self.h, self._shape_array(shapes), u32(len(shapes)), C.c_int(1 if as_set else 0),
                C.c_int({"grid": 0, "bcc": 1}[lattice]), f64(spacing), (f64 * 3)(*initial_velocity), C.byref(n))
        return int(n.value)

    def set_collider_shapes(self, shapes, as_set=True, linear_velocity=(0.0, 0.0, 0.0), angular_velocity=(0.0, 0.0, 0.0),
                            friction=0.0):
        """RigidBodyCollider3 over analytic surfaces with identity transforms.  `shapes`: dicts with
        kind "box" (a = lower, b = upper corner) | "sphere" (a = centre, b[0] = radius) | "cylinder" (a = centre,
        b[0] = radius, b[1] = height, axis y) | "plane" (a = point, b = unit normal), optional flipped=True
        (Surface3::isNormalFlipped).  as_set: the collider's surface is an ImplicitSurfaceSet3 of them (required for more
        than one); otherwise the single surface itself.  An empty list removes the collider."""
        self._c("hybrid_set_collider_shapes", self.h, self._shape_array(shapes), u32(len(shapes)), C.c_int(1 if as_set else 0),
                (f64 * 3)(*linear_velocity), (f64 * 3)(*angular_velocity), f64(friction))

    def set_collider_shape(self, kind, a=(0.0, 0.0, 0.0), b=(0.0, 0.0, 0.0), linear_velocity=(0.0, 0.0, 0.0),
                           angular_velocity=(0.0, 0.0, 0.0), friction=0.0, flipped=False):
        """One analytic surface as the collider (see set_collider_shapes); None / "none" removes it."""
        shapes = [] if kind in (None, "none") else [dict(kind=kind, a=a, b=b, flipped=flipped)]
        self.set_collider_shapes(shapes, False, linear_velocity, angular_velocity, friction)

    def get_collider_sdf(self):
        s = np.zeros(size3(self.grid.res))
        self._c("hybrid_get_collider_sdf", self.h, _p(s))
        return s

    def set_velocity(self, u, v, w):
        self._c("hybrid_set_velocity", self.h, _p(c64(u)), _p(c64(v)), _p(c64(w)))

    def get_velocity(self):
        g = self.grid
        u = np.zeros(size3(g.usize))
        v = np.zeros(size3(g.vsize))
        w = np.zeros(size3(g.wsize))
        self._c("hybrid_get_velocity", self.h, _p(u), _p(v), _p(w))
        return u, v, w

    def get_sdf(self):
        s = np.zeros(size3(self.grid.res))
        self._c("hybrid_get_sdf", self.h, _p(s))
        return s

    def get_pressure(self):
        s = np.zeros(size3(self.grid.res))
        self._c("hybrid_get_pressure", self.h, _p(s))
        return s

    def get_markers(self):
        g = self.grid
        um = np.zeros(size3(g.usize), dtype=np.int8)
        vm = np.zeros(size3(g.vsize), dtype=np.int8)
        wm = np.zeros(size3(g.wsize), dtype=np.int8)
        pc = C.POINTER(C.c_char)
        self._c("hybrid_get_markers", self.h, _p(um, pc), _p(vm, pc), _p(wm, pc))
        return um, vm, wm

    def cfl(self, dt):
        r = f64(0)
        self._c("hybrid_cfl", self.h, f64(dt), C.byref(r))
        return r.value

    def num_sub_steps(self, dt):
        r = u32(0)
        self._c("hybrid_num_sub_steps", self.h, f64(dt), C.byref(r))
        return r.value

    def step(self, dt):
        st = StepStats()
        self._c("hybrid_step", self.h, f64(dt), C.byref(st))
        return st

    def advance_frame(self, dt):
        st = StepStats()
        n = u32(0)
        self._c("hybrid_advance_frame", self.h, f64(dt), C.byref(n), C.byref(st))
        return n.value, st

    def apply_boundary_condition(self):
        self._c("hybrid_apply_boundary_condition", self.h)

    def transfer_particles_to_grids(self):
        self._c("hybrid_transfer_particles_to_grids", self.h)

    def build_signed_distance_field(self):
        self._c("hybrid_build_signed_distance_field", self.h)

    def extrapolate_velocity_to_air(self):
        self._c("hybrid_extrapolate_velocity_to_air", self.h)

    def begin_advance_time_step(self, dt):
        self._c("hybrid_begin_advance_time_step", self.h, f64(dt))

    def compute_external_forces(self, dt):
        self._c("hybrid_compute_external_forces", self.h, f64(dt))

    def compute_pressure(self, dt):
        r = MgpcgResult()
        self._c("hybrid_compute_pressure", self.h, f64(dt), C.byref(r))
        return r

    def set_viscosity(self, coefficient):
        """GridFluidSolver3::setViscosityCoefficient: > kEpsilonD switches computeViscosity on in step()."""
        self._c("hybrid_set_viscosity", self.h, f64(coefficient))

    def compute_viscosity(self, dt):
        r = (MgpcgResult * 3)()
        self._c("hybrid_compute_viscosity", self.h, f64(dt), r)
        return list(r)

    def transfer_grids_to_particles(self):
        self._c("hybrid_transfer_grids_to_particles", self.h)

    def move_particles(self, dt):
        self._c("hybrid_move_particles", self.h, f64(dt))

    def compute_advection(self, dt):
        self._c("hybrid_compute_advection", self.h, f64(dt))



class JetcuError(RuntimeError):
    def __init__(self, name, rc, msg):
        super().__init__("jetcu_%s failed (rc=%d): %s" % (name, rc, msg))
        self.rc = rc


class Jetcu(LinearApi):
    """One jetcu context (one GPU).  Methods mirror the reference's operator names; see include/jetcu.h."""

    def __init__(self, device=0, lib_path=LIB_PATH):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(
                "%s not built: run `python fluid-engine-dev_b200/build.py` (there is no CPU fallback)" % lib_path)
        self.lib = C.CDLL(lib_path)
        self.lib.jetcu_last_error.restype = C.c_char_p
        self.lib.jetcu_last_error.argtypes = [C.c_void_p]
        self.lib.jetcu_version.restype = C.c_char_p
        self.lib.jetcu_launch_count.restype = u64
        self.lib.jetcu_launch_count.argtypes = [C.c_void_p]
        self.lib.jetcu_stream.restype = C.c_void_p
        self.lib.jetcu_stream.argtypes = [C.c_void_p]
        self.lib.jetcu_mg_num_levels.restype = u32
        self.lib.jetcu_mg_num_levels.argtypes = [u64, u64, u64, u32]
        for name in ("jetcu_hybrid_num_particles", "jetcu_hybrid_num_local_particles"):
            if hasattr(self.lib, name):
                getattr(self.lib, name).restype = u64
                getattr(self.lib, name).argtypes = [C.c_void_p]
        self.ctx = C.c_void_p()
        rc = self.lib.jetcu_ctx_create(int(device), C.byref(self.ctx))
        if rc != 0:
            raise JetcuError("ctx_create", rc, self.lib.jetcu_last_error(None).decode())
        self.kind = "jetcu"
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.jetcu_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def version(self):
        return self.lib.jetcu_version().decode()

    def launch_count(self):
        return int(self.lib.jetcu_launch_count(self.ctx))

    def stream(self):
        return self.lib.jetcu_stream(self.ctx)

    def synchronize(self):
        self._call_raw("synchronize", self.ctx)

    def _fn(self, name, raw=False):
        return getattr(self.lib, "jetcu_" + name)

    def _call_raw(self, name, *args, raw=False):
        fn = getattr(self.lib, "jetcu_" + name)
        if raw:
            return fn(*args)
        fn.restype = C.c_int
        rc = fn(*args)
        if rc != 0:
            raise JetcuError(name, rc, self.lib.jetcu_last_error(self.ctx).decode())
        return rc

    def _call(self, name, *args, raw=False):
        # stateless entry points take the ctx first
        return self._call_raw(name, self.ctx, *args, raw=raw)

    def hybrid(self, params):
        h = C.c_void_p()
        self._call("hybrid_create", C.byref(params), C.byref(h))
        return HybridApi(self._call_raw, h, params)

    def smoke(self, params):
        h = C.c_void_p()
        self._call("smoke_create", C.byref(params), C.byref(h))
        return SmokeApi(self._call_raw, h, params)

    # ------------------------------------------------------------------ z-slab communicator (include/jetcu.h)
    def comm_unique_id(self):
        """128-byte NCCL id; call on rank 0 and hand it to the other ranks (e.g. torch.distributed.broadcast)."""
        buf = (C.c_uint8 * 128)()
        self._call_raw("comm_unique_id", buf)
        return bytes(buf)

    def comm_init(self, world, rank, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._call("comm_init", C.c_int(world), C.c_int(rank), buf)
        self.world, self.rank = world, rank

    def comm_init_torch(self):
        """Joins the z-slab communicator of the running torch.distributed job (one rank per GPU)."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        if world == 1:
            return
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        self.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))

    def comm_stats(self):
        m, b = u64(0), u64(0)
        self._call("comm_stats", C.byref(m), C.byref(b))
        return int(m.value), int(b.value)

    def comm_p2p_enabled(self):
        """True when halo planes travel through peer memory (JETCU_P2P_HALO=1) instead of NCCL send/recv."""
        return bool(self._call("comm_p2p_enabled", raw=True))

    def comm_halo_benchmark(self, plane_bytes, depth=1, iters=200):
        """Average device milliseconds per halo exchange (both neighbours) for the active transport."""
        ms = f64(0.0)
        self._call("comm_halo_benchmark", u64(plane_bytes), C.c_int(depth), C.c_int(iters), C.byref(ms))
        return float(ms.value)

    def slab_range(self, nz, world, rank):
        z0, z1 = u64(0), u64(0)
        self._call_raw("slab_range", u64(nz), C.c_int(world), C.c_int(rank), C.byref(z0), C.byref(z1))
        return int(z0.value), int(z1.value)
