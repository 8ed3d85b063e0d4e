// diffusion_ops.cuh — per-sample bodies of GridBackwardEulerDiffusionSolver3 (src/jet/grid_backward_euler_diffusion_solver3.cpp)
// as __host__ __device__ functions: diffusion.cu's kernels are thin wrappers, tests/hostcheck runs the same source on the
// CPU against the oracle (no reductions here: bit for bit).  SURVEY §8f-4.
#pragma once
#include "collider_ops.cuh"

#define DIFF_FLUID 0      // kFluid, kAir, kBoundary — :16-18
#define DIFF_AIR 1
#define DIFF_BOUNDARY 2

// One sample array of a solve: its extent and the position of sample (0, 0, 0) (uPosition / vPosition / wPosition /
// dataPosition of the grid: dataOrigin + h * (i, j, k)).
struct DiffArray {
    int sx, sy, sz;
    double dox, doy, doz;
};

// kind 0 = cell-centred scalar grid, 1/2/3 = u/v/w of a face-centred grid
__host__ __device__ __forceinline__ DiffArray diff_array(const GridD& g, int kind) {
    DiffArray a;
    a.sx = g.nx + (kind == 1); a.sy = g.ny + (kind == 2); a.sz = g.nz + (kind == 3);
    a.dox = g.ox + 0.5 * (kind == 1 ? 0.0 : g.hx);
    a.doy = g.oy + 0.5 * (kind == 2 ? 0.0 : g.hy);
    a.doz = g.oz + 0.5 * (kind == 3 ? 0.0 : g.hz);
    return a;
}

// buildMarkers, :166-183: boundary first (isInsideSdf(phi) = phi < 0, level_set_utils-inl.h:19-22), then fluid, else air.
// csdf / fsdf: cell-centred SDF arrays sampled with the linear sampler (ScalarGrid3::sample), or nullptr for the constant
// fields the fluid solvers pass when there is none (collider SDF grid filled with kMaxD; ConstantScalarField3(-kMaxD)).
__host__ __device__ __forceinline__ char diff_marker(const GridD& g, const DiffArray& a, const double* csdf, const double* fsdf,
                                                     int i, int j, int k) {
    const double px = a.dox + g.hx * (double)i, py = a.doy + g.hy * (double)j, pz = a.doz + g.hz * (double)k;
    if (csdf && jc_sample(collider_view(g, csdf), px, py, pz) < 0.0) return DIFF_BOUNDARY;
    if (!fsdf || jc_sample(collider_view(g, fsdf), px, py, pz) < 0.0) return DIFF_FLUID;
    return DIFF_AIR;
}

// buildMatrix, :185-251: one row (center, right, up, front)
__host__ __device__ __forceinline__ void diff_matrix_row(const char* mk, int sx, int sy, int sz, int i, int j, int k, double cx,
                                                         double cy, double cz, bool dirichlet, double* center, double* right,
                                                         double* up, double* front) {
    const size_t py = (size_t)sx, pz = (size_t)sx * sy, c = (size_t)k * pz + (size_t)j * py + i;
    double ce = 1.0, ri = 0.0, u = 0.0, fr = 0.0;
    if (mk[c] == DIFF_FLUID) {
        if (i + 1 < sx) {
            if ((dirichlet && mk[c + 1] != DIFF_AIR) || mk[c + 1] == DIFF_FLUID) ce += cx;
            if (mk[c + 1] == DIFF_FLUID) ri -= cx;
        }
        if (i > 0 && ((dirichlet && mk[c - 1] != DIFF_AIR) || mk[c - 1] == DIFF_FLUID)) ce += cx;
        if (j + 1 < sy) {
            if ((dirichlet && mk[c + py] != DIFF_AIR) || mk[c + py] == DIFF_FLUID) ce += cy;
            if (mk[c + py] == DIFF_FLUID) u -= cy;
        }
        if (j > 0 && ((dirichlet && mk[c - py] != DIFF_AIR) || mk[c - py] == DIFF_FLUID)) ce += cy;
        if (k + 1 < sz) {
            if ((dirichlet && mk[c + pz] != DIFF_AIR) || mk[c + pz] == DIFF_FLUID) ce += cz;
            if (mk[c + pz] == DIFF_FLUID) fr -= cz;
        }
        if (k > 0 && ((dirichlet && mk[c - pz] != DIFF_AIR) || mk[c - pz] == DIFF_FLUID)) ce += cz;
    }
    *center = ce; *right = ri; *up = u; *front = fr;
}

// buildVectors, :253-291: b (the solver zeroes x itself, fdm_iccg_solver3.cpp:169)
__host__ __device__ __forceinline__ double diff_rhs(const char* mk, const double* f, int sx, int sy, int sz, int i, int j, int k,
                                                    double cx, double cy, double cz, bool dirichlet) {
    const size_t py = (size_t)sx, pz = (size_t)sx * sy, c = (size_t)k * pz + (size_t)j * py + i;
    double v = f[c];
    if (dirichlet && mk[c] == DIFF_FLUID) {
        if (i + 1 < sx && mk[c + 1] == DIFF_BOUNDARY) v += cx * f[c + 1];
        if (i > 0 && mk[c - 1] == DIFF_BOUNDARY) v += cx * f[c - 1];
        if (j + 1 < sy && mk[c + py] == DIFF_BOUNDARY) v += cy * f[c + py];
        if (j > 0 && mk[c - py] == DIFF_BOUNDARY) v += cy * f[c - py];
        if (k + 1 < sz && mk[c + pz] == DIFF_BOUNDARY) v += cz * f[c + pz];
        if (k > 0 && mk[c - pz] == DIFF_BOUNDARY) v += cz * f[c - pz];
    }
    return v;
}
