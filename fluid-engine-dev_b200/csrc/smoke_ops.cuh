// smoke_ops.cuh — per-cell / per-face bodies of the GridSmokeSolver3 kernels (BASELINE config 5) as
// __host__ __device__ functions: smoke.cu's kernels are thin wrappers, tests/hostcheck runs the same source on the CPU
// against the oracle (oracle/jet_port_smoke.inc, itself bit-identical to the unmodified reference).
//   buoyancy            GridSmokeSolver3::computeBuoyancyForce      src/jet/grid_smoke_solver3.cpp:131-187
//   advection           SemiLagrangian3::advect / backTrace         src/jet/semi_lagrangian3.cpp:21-160
//                       CubicSemiLagrangian3 samplers               src/jet/cubic_semi_lagrangian3.cpp:15-55
//                       CubicArraySampler3::operator()              include/jet/detail/array_samplers3-inl.h:290-346
//                       monotonicCatmullRom                         include/jet/detail/math_utils-inl.h:233-261
// No collider: the boundary SDF is the collider grid filled with kMaxD, so `boundarySdf.sample(x) > 0` always holds and
// backTrace's boundary branch (phi0 * phi1 < 0) can never trigger; both are left out.
#pragma once
#include "collider_ops.cuh"

#define SM_MAX_BACKTRACE_STEPS 100000

__host__ __device__ __forceinline__ double sm_sign(double x) { return x >= 0 ? 1.0 : -1.0; }  // math_utils-inl.h:23-29

__host__ __device__ __forceinline__ double sm_catmull_rom(double f0, double f1, double f2, double f3, double f) {
    double d1 = (f2 - f0) / 2;
    double d2 = (f3 - f1) / 2;
    const double D1 = f2 - f1;
    if (fabs(D1) < 2.220446049250313e-16) { d1 = d2 = 0; }  // kEpsilonD
    if (sm_sign(D1) != sm_sign(d1)) d1 = 0;
    if (sm_sign(D1) != sm_sign(d2)) d2 = 0;
    const double a3 = d1 + d2 - 2 * D1;
    const double a2 = 3 * D1 - 2 * d1 - d2;
    const double a1 = d1;
    const double a0 = f1;
    return a3 * (f * f * f) + a2 * (f * f) + a1 * f + a0;
}

// CubicArraySampler3<double,double>::operator(): normalisation with the DIVISION by h (jc_norm), 4x4x4 clamped
// neighbourhood, Catmull-Rom along x, then y, then z.
__host__ __device__ __forceinline__ double sm_sample_cubic(const GridView& g, double px, double py, double pz) {
    int i, j, k;
    double fx, fy, fz;
    jc_barycentric(jc_norm(px, g.ox, g.hx, g.ihx), g.nx - 1, &i, &fx);
    jc_barycentric(jc_norm(py, g.oy, g.hy, g.ihy), g.ny - 1, &j, &fy);
    jc_barycentric(jc_norm(pz, g.oz, g.hz, g.ihz), g.nz - 1, &k, &fz);
    const int is[4] = {i - 1 > 0 ? i - 1 : 0, i, i + 1 < g.nx - 1 ? i + 1 : g.nx - 1, i + 2 < g.nx - 1 ? i + 2 : g.nx - 1};
    const int js[4] = {j - 1 > 0 ? j - 1 : 0, j, j + 1 < g.ny - 1 ? j + 1 : g.ny - 1, j + 2 < g.ny - 1 ? j + 2 : g.ny - 1};
    const int ks[4] = {k - 1 > 0 ? k - 1 : 0, k, k + 1 < g.nz - 1 ? k + 1 : g.nz - 1, k + 2 < g.nz - 1 ? k + 2 : g.nz - 1};
    const size_t sy = (size_t)g.nx, sz = (size_t)g.nx * g.ny;
    const double* d = g.data;
    double kv[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        double jv[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const size_t row = (size_t)ks[kk] * sz + (size_t)js[jj] * sy;
            jv[jj] = sm_catmull_rom(d[row + is[0]], d[row + is[1]], d[row + is[2]], d[row + is[3]], fx);
        }
        kv[kk] = sm_catmull_rom(jv[0], jv[1], jv[2], jv[3], fy);
    }
    return sm_catmull_rom(kv[0], kv[1], kv[2], kv[3], fz);
}

// FaceCenteredGrid3::sample with the three LINEAR component samplers (the flow field of backTrace)
__host__ __device__ __forceinline__ void sm_flow(const GridD& g, const double* u, const double* v, const double* w, const double* p,
                                                 double* out) {
    out[0] = jc_sample(comp_view(g, 0, u), p[0], p[1], p[2]);
    out[1] = jc_sample(comp_view(g, 1, v), p[0], p[1], p[2]);
    out[2] = jc_sample(comp_view(g, 2, w), p[0], p[1], p[2]);
}

// SemiLagrangian3::backTrace — src/jet/semi_lagrangian3.cpp:117-160 (adaptive mid-point sub-steps)
// `csdf` = the collider SDF grid (boundarySdf = *colliderSdf(), sampled linearly), nullptr for a solver without collider: its
// grid holds kMaxD, phi0 * phi1 is +inf and the boundary branch cannot trigger.
__host__ __device__ __forceinline__ GridView sm_cell_view(const GridD& g, const double* data);
__host__ __device__ __forceinline__ void sm_back_trace(const GridD& g, const double* u, const double* v, const double* w, double dt,
                                                       double h, const double* start, double* out, const double* csdf = nullptr) {
    double remaining = dt;
    double pt0[3] = {start[0], start[1], start[2]}, pt1[3] = {start[0], start[1], start[2]};
    // The reference loops until the time is used up; the number of sub-steps is ~ |v| dt / h, i.e. a handful in any sane
    // flow.  When the (fragile) reference pressure solver has diverged, |v| can reach 1e8 and the loop runs
    // for minutes per point on the CPU; on the GPU that would look like a hang, so the trace gives up after
    // SM_MAX_BACKTRACE_STEPS sub-steps (the field is garbage by then in the reference as well).
    for (int guard = 0; remaining > 2.220446049250313e-16 && guard < SM_MAX_BACKTRACE_STEPS; ++guard) {
        double vel0[3], mv[3], mid[3];
        sm_flow(g, u, v, w, pt0, vel0);
        const double len = sqrt(vel0[0] * vel0[0] + vel0[1] * vel0[1] + vel0[2] * vel0[2]);
        const double nsub = cd_max(ceil(len * remaining / h), 1.0);
        dt = remaining / nsub;
        const double hdt = 0.5 * dt;
#pragma unroll
        for (int a = 0; a < 3; ++a) mid[a] = pt0[a] - vel0[a] * hdt;
        sm_flow(g, u, v, w, mid, mv);
#pragma unroll
        for (int a = 0; a < 3; ++a) pt1[a] = pt0[a] - mv[a] * dt;
        if (csdf) {  // boundary handling (:146-153): stop at the collider surface
            const GridView bs = sm_cell_view(g, csdf);
            const double phi0 = jc_sample(bs, pt0[0], pt0[1], pt0[2]);
            const double phi1 = jc_sample(bs, pt1[0], pt1[1], pt1[2]);
            if (phi0 * phi1 < 0.0) {
                const double wgt = fabs(phi1) / (fabs(phi0) + fabs(phi1));
#pragma unroll
                for (int a = 0; a < 3; ++a) pt1[a] = pt0[a] * wgt + pt1[a] * (1.0 - wgt);  // w * pt0 + (1.0 - w) * pt1
                break;
            }
        }
        remaining -= dt;
        pt0[0] = pt1[0]; pt0[1] = pt1[1]; pt0[2] = pt1[2];
    }
    out[0] = pt1[0]; out[1] = pt1[1]; out[2] = pt1[2];
}

__host__ __device__ __forceinline__ GridView sm_cell_view(const GridD& g, const double* data) {
    GridView v;
    v.data = data; v.nx = g.nx; v.ny = g.ny; v.nz = g.nz; v.hx = g.hx; v.hy = g.hy; v.hz = g.hz;
    v.ox = g.ox + 0.5 * g.hx; v.oy = g.oy + 0.5 * g.hy; v.oz = g.oz + 0.5 * g.hz;
    v.ihx = g.ihx; v.ihy = g.ihy; v.ihz = g.ihz;
    return v;
}

// SemiLagrangian3::advect(ScalarGrid3): one cell of a cell-centred field (in = the clone, flow = current velocity)
// `linear`: SemiLagrangian3 itself (input.sampler(), the trilinear LinearArraySampler3 — semi_lagrangian3.cpp:162-176) instead of
// CubicSemiLagrangian3's overrides (cubic_semi_lagrangian3.cpp:15-55)
__host__ __device__ __forceinline__ double sm_advect_cell(const GridD& g, const double* in, const double* u, const double* v,
                                                          const double* w, double dt, double hmin, int i, int j, int k,
                                                          bool linear = false, const double* csdf = nullptr) {
    const GridView si = sm_cell_view(g, in);
    const double start[3] = {si.ox + g.hx * (double)i, si.oy + g.hy * (double)j, si.oz + g.hz * (double)k};
    // inside the collider the output grid is not written (semi_lagrangian3.cpp:36-40): it keeps its value, which `in` cloned
    if (csdf && !(jc_sample(sm_cell_view(g, csdf), start[0], start[1], start[2]) > 0.0)) return in[((size_t)k * g.ny + j) * g.nx + i];
    double pt[3];
    sm_back_trace(g, u, v, w, dt, hmin, start, pt, csdf);
    return linear ? jc_sample(si, pt[0], pt[1], pt[2]) : sm_sample_cubic(si, pt[0], pt[1], pt[2]);
}

// SemiLagrangian3::advect(FaceCenteredGrid3): one face of component AXIS (input = flow = the velocity clone u0, v0, w0)
template <int AXIS>
__host__ __device__ __forceinline__ double sm_advect_face(const GridD& g, const double* u0, const double* v0, const double* w0,
                                                          double dt, double hmin, int i, int j, int k, bool linear = false,
                                                          const double* csdf = nullptr) {
    const GridView sc = comp_view(g, AXIS, AXIS == 0 ? u0 : (AXIS == 1 ? v0 : w0));
    const double start[3] = {sc.ox + g.hx * (double)i, sc.oy + g.hy * (double)j, sc.oz + g.hz * (double)k};
    if (csdf && !(jc_sample(sm_cell_view(g, csdf), start[0], start[1], start[2]) > 0.0))   // :88-118, per component
        return sc.data[((size_t)k * (g.ny + (AXIS == 1)) + j) * (g.nx + (AXIS == 0)) + i];
    double pt[3];
    sm_back_trace(g, u0, v0, w0, dt, hmin, start, pt, csdf);
    return linear ? jc_sample(sc, pt[0], pt[1], pt[2]) : sm_sample_cubic(sc, pt[0], pt[1], pt[2]);
}

// computeBuoyancyForce: one face of component AXIS; up_a = component AXIS of -gravity.normalized(), t_amb = mean temperature
template <int AXIS>
__host__ __device__ __forceinline__ void sm_buoyancy_face(const GridD& g, const double* den, const double* temp, double fd, double ft,
                                                          double t_amb, double dt, double up_a, double* comp, int i, int j, int k) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1);
    double pt[3];
    face_position<AXIS>(g, i, j, k, pt);
    const double fBuoy = fd * jc_sample(sm_cell_view(g, den), pt[0], pt[1], pt[2]) +
                         ft * (jc_sample(sm_cell_view(g, temp), pt[0], pt[1], pt[2]) - t_amb);
    comp[((size_t)k * sy + j) * sx + i] += dt * fBuoy * up_a;
}

// VolumeGridEmitter3 with a step-function target (src/jet/volume_grid_emitter3.cpp:33-42, 83-99), one cell: the source of
// src/examples/smoke_sim.  smearedHeavisideSdf — include/jet/detail/level_set_utils-inl.h:29-40 (float literals 0.5f as there).
__host__ __device__ __forceinline__ double sm_smeared_heaviside_sdf(double phi) {
    if (phi > 1.5) return 1;
    if (phi < -1.5) return 0;
    return 0.5f + phi / 3.0 + 0.5f * (1.0 / 3.14159265358979323846264338327950288) * sin(3.14159265358979323846264338327950288 * phi / 1.5);
}
__host__ __device__ __forceinline__ void sm_emit_step_cell(const GridD& g, const ColliderD& e, double width, double lo, double hi,
                                                           double* field, int i, int j, int k) {
    const double sdf = collider_sdf_cell(g, e, i, j, k);   // sourceRegion()->signedDistance(dataPosition(i, j, k))
    const double step = 1.0 - sm_smeared_heaviside_sdf(sdf / width);
    double* cell = field + ((size_t)k * g.ny + j) * g.nx + i;
    *cell = cd_max(*cell, (hi - lo) * step + lo);
}

// GridFluidSolver3::cfl, one cell (linear index c): the largest SIGNED component of valueAtCellCenter + dt * gravity
// (src/jet/grid_fluid_solver3.cpp:60-66, face_centered_grid3.cpp:128-135); dtg = dt * gravity
__host__ __device__ __forceinline__ double sm_cfl_cell(const GridD& g, const double* u, const double* v, const double* w, double dtgx,
                                                       double dtgy, double dtgz, size_t c) {
    const int i = (int)(c % g.nx);
    const size_t t = c / g.nx;
    const int j = (int)(t % g.ny), k = (int)(t / g.ny);
    const size_t iu = ((size_t)k * g.ny + j) * (g.nx + 1) + i;
    const size_t iv = ((size_t)k * (g.ny + 1) + j) * g.nx + i;
    const double vx = 0.5 * (u[iu] + u[iu + 1]) + dtgx;
    const double vy = 0.5 * (v[iv] + v[iv + g.nx]) + dtgy;
    const double vz = 0.5 * (w[c] + w[c + (size_t)g.nx * g.ny]) + dtgz;
    double m = 0.0;
    m = fmax(m, vx); m = fmax(m, vy); m = fmax(m, vz);
    return m;
}

