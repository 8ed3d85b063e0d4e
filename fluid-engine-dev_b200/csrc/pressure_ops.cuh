// pressure_ops.cuh — one-cell / one-face bodies of the pressure-system kernels (GridFractionalSinglePhasePressureSolver3:
// buildWeights, restrict, buildSingleSystem, applyPressureGradient) as __host__ __device__ functions.  pressure.cu's kernels are
// thin wrappers; tests/hostcheck runs the same source on the CPU against the oracle, bit for bit.
// Reference: src/jet/grid_fractional_single_phase_pressure_solver3.cpp (line ranges cited per function).
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "sampling.cuh"

struct GridP {
    int nx, ny, nz;
    double hx, hy, hz, ox, oy, oz;
};

// _fluidSdf[0](i,j,k) = float(fluidSdf.sample(cellCenterPosition(i,j,k)))            (:546-549)
// cellCenter = o + h*(i+0.5) (src/jet/grid3.cpp:32-38); cell-centred data origin = o + 0.5*h
// (src/jet/cell_centered_scalar_grid3.cpp:60-62).  fluid == nullptr => ConstantScalarField3(-kMaxD).
__host__ __device__ __forceinline__ void fluid_sdf0_cell(const GridP& g, const double* __restrict__ fluid, float* __restrict__ out,
                                                         int i, int j, int k) {
    double val;
    if (fluid) {
        GridView v{fluid, g.nx, g.ny, g.nz, g.hx, g.hy, g.hz, g.ox + 0.5 * g.hx, g.oy + 0.5 * g.hy, g.oz + 0.5 * g.hz};
        val = jc_sample(v, g.ox + g.hx * (i + 0.5), g.oy + g.hy * (j + 0.5), g.oz + g.hz * (k + 0.5));
    } else {
        val = -DBL_MAX;
    }
    out[((size_t)k * g.ny + j) * g.nx + i] = (float)val;
}

__host__ __device__ __forceinline__ float weight_from_frac(double frac) {
    // clamp(1 - frac, 0, 1), values in (0, kMinWeight) raised to kMinWeight = 0.01   (:563-571)
    double weight = jc_clamp(1.0 - frac, 0.0, 1.0);
    if (weight < 0.01 && weight > 0.0) weight = 0.01;
    return (float)weight;
}

// _{u,v,w}Weights[0]: open fraction of each face from the collider SDF at its 4 corners (:551-617).
// AXIS = 0/1/2 for u/v/w.  Face position = faceOrigin + h*(i,j,k) (src/jet/face_centered_grid3.cpp:216-238,442-444).
template <int AXIS>
__host__ __device__ __forceinline__ void face_weight0_face(const GridP& g, const double* __restrict__ collider, float* __restrict__ out,
                                                           int i, int j, int k) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1);
    size_t f = ((size_t)k * sy + j) * sx + i;
    if (!collider) { out[f] = 1.0f; return; }  // boundarySdf == +kMaxD: fractionInside = 0 -> weight 1
    GridView v{collider, g.nx, g.ny, g.nz, g.hx, g.hy, g.hz, g.ox + 0.5 * g.hx, g.oy + 0.5 * g.hy, g.oz + 0.5 * g.hz};
    double fox = g.ox + 0.5 * (AXIS == 0 ? 0.0 : g.hx);
    double foy = g.oy + 0.5 * (AXIS == 1 ? 0.0 : g.hy);
    double foz = g.oz + 0.5 * (AXIS == 2 ? 0.0 : g.hz);
    double px = fox + g.hx * i, py = foy + g.hy * j, pz = foz + g.hz * k;
    double p0, p1, p2, p3;
    if (AXIS == 0) {
        p0 = jc_sample(v, px + 0.0, py + -0.5 * g.hy, pz + -0.5 * g.hz);
        p1 = jc_sample(v, px + 0.0, py + 0.5 * g.hy, pz + -0.5 * g.hz);
        p2 = jc_sample(v, px + 0.0, py + -0.5 * g.hy, pz + 0.5 * g.hz);
        p3 = jc_sample(v, px + 0.0, py + 0.5 * g.hy, pz + 0.5 * g.hz);
    } else if (AXIS == 1) {
        p0 = jc_sample(v, px + -0.5 * g.hx, py + 0.0, pz + -0.5 * g.hz);
        p1 = jc_sample(v, px + -0.5 * g.hx, py + 0.0, pz + 0.5 * g.hz);
        p2 = jc_sample(v, px + 0.5 * g.hx, py + 0.0, pz + -0.5 * g.hz);
        p3 = jc_sample(v, px + 0.5 * g.hx, py + 0.0, pz + 0.5 * g.hz);
    } else {
        p0 = jc_sample(v, px + -0.5 * g.hx, py + -0.5 * g.hy, pz + 0.0);
        p1 = jc_sample(v, px + -0.5 * g.hx, py + 0.5 * g.hy, pz + 0.0);
        p2 = jc_sample(v, px + 0.5 * g.hx, py + -0.5 * g.hy, pz + 0.0);
        p3 = jc_sample(v, px + 0.5 * g.hx, py + 0.5 * g.hy, pz + 0.0);
    }
    out[f] = weight_from_frac(jc_frac_inside4(p0, p1, p2, p3));
}

// Anonymous-namespace restrict(Array3<float>) (:29-115): per axis either the centred (1,3,3,1)/8 4-tap kernel or,
// where finer != 2*coarser (the face grid's staggered axis), the 3-tap {0,1,0} kernel; float accumulation z,y,x.
__host__ __device__ __forceinline__ void restrict_f32_cell(const float* __restrict__ finer, int fx, int fy, int fz,
                                                           float* __restrict__ coarser, int cx, int cy, int cz, int i, int j, int k) {
    const float centered[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    const float staggered[4] = {0.f, 1.f, 0.f, 0.f};
    const int ksx = (fx != 2 * cx) ? 3 : 4, ksy = (fy != 2 * cy) ? 3 : 4, ksz = (fz != 2 * cz) ? 3 : 4;
    int iI[4], jI[4], kI[4];
    iI[0] = (i > 0) ? 2 * i - 1 : 2 * i; iI[1] = 2 * i;
    if (ksx == 3) { iI[2] = (i + 1 < cx) ? 2 * i + 1 : 2 * i; iI[3] = 0; }
    else { iI[2] = 2 * i + 1; iI[3] = (i + 1 < cx) ? 2 * i + 2 : 2 * i + 1; }
    jI[0] = (j > 0) ? 2 * j - 1 : 2 * j; jI[1] = 2 * j;
    if (ksy == 3) { jI[2] = (j + 1 < cy) ? 2 * j + 1 : 2 * j; jI[3] = 0; }
    else { jI[2] = 2 * j + 1; jI[3] = (j + 1 < cy) ? 2 * j + 2 : 2 * j + 1; }
    kI[0] = (k > 0) ? 2 * k - 1 : 2 * k; kI[1] = 2 * k;
    if (ksz == 3) { kI[2] = (k + 1 < cz) ? 2 * k + 1 : 2 * k; kI[3] = 0; }
    else { kI[2] = 2 * k + 1; kI[3] = (k + 1 < cz) ? 2 * k + 2 : 2 * k + 1; }
    float sum = 0.0f;
    for (int z = 0; z < ksz; ++z)
        for (int y = 0; y < ksy; ++y)
            for (int x = 0; x < ksx; ++x) {
                float kx = (ksx == 3) ? staggered[x] : centered[x];
                float ky = (ksy == 3) ? staggered[y] : centered[y];
                float kz = (ksz == 3) ? staggered[z] : centered[z];
                float w = kx * ky * kz;
                sum += w * finer[((size_t)kI[z] * fy + jI[y]) * fx + iI[x]];
            }
    coarser[((size_t)k * cy + j) * cx + i] = sum;
}

// buildSingleSystem, uncompressed (:117-267).  Writes the SoA matrix; b only on the finest level (b == nullptr on
// coarse levels: it is dead there, the V-cycle overwrites b[l>0] with the restricted residual — mg-inl.h:30).
__host__ __device__ __forceinline__ void build_system_cell(
    int nx, int ny, int nz, double hx, double hy, double hz, const float* __restrict__ sdf, const float* __restrict__ uw,
    const float* __restrict__ vw, const float* __restrict__ ww, const double* __restrict__ u, const double* __restrict__ v,
    const double* __restrict__ w, const double* __restrict__ bu, const double* __restrict__ bv, const double* __restrict__ bw,
    double* __restrict__ Acenter, float* __restrict__ Acr, float* __restrict__ Acu, float* __restrict__ Acf,
    uint8_t* __restrict__ Acode, double* __restrict__ b, int i, int j, int k) {
    const double ihx = 1.0 / hx, ihy = 1.0 / hy, ihz = 1.0 / hz;
    const double ihx2 = ihx * ihx, ihy2 = ihy * ihy, ihz2 = ihz * ihz;
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    size_t c = (size_t)k * sz + (size_t)j * sy + i;
    size_t u0 = ((size_t)k * ny + j) * (nx + 1) + i, u1 = u0 + 1;
    size_t v0 = ((size_t)k * (ny + 1) + j) * nx + i, v1 = v0 + nx;
    size_t w0 = c, w1 = c + sz;
    double center = 0.0, bb = 0.0;
    float right = 0.f, up = 0.f, front = 0.f;  // packed off-diagonals: the face weight where `row.right -= term` applies
    double cphi = sdf[c];
    if (cphi < 0) {
        double term, theta;
        if (i + 1 < nx) {
            term = uw[u1] * ihx2;
            double nphi = sdf[c + 1];
            if (nphi < 0) { center += term; right = uw[u1]; }
            else { theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01); center += term / theta; }
            if (b) bb += uw[u1] * u[u1] * ihx;
        } else if (b) { bb += u[u1] * ihx; }
        if (i > 0) {
            term = uw[u0] * ihx2;
            double nphi = sdf[c - 1];
            if (nphi < 0) { center += term; }
            else { theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01); center += term / theta; }
            if (b) bb -= uw[u0] * u[u0] * ihx;
        } else if (b) { bb -= u[u0] * ihx; }
        if (j + 1 < ny) {
            term = vw[v1] * ihy2;
            double nphi = sdf[c + sy];
            if (nphi < 0) { center += term; up = vw[v1]; }
            else { theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01); center += term / theta; }
            if (b) bb += vw[v1] * v[v1] * ihy;
        } else if (b) { bb += v[v1] * ihy; }
        if (j > 0) {
            term = vw[v0] * ihy2;
            double nphi = sdf[c - sy];
            if (nphi < 0) { center += term; }
            else { theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01); center += term / theta; }
            if (b) bb -= vw[v0] * v[v0] * ihy;
        } else if (b) { bb -= v[v0] * ihy; }
        if (k + 1 < nz) {
            term = ww[w1] * ihz2;
            double nphi = sdf[c + sz];
            if (nphi < 0) { center += term; front = ww[w1]; }
            else { theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01); center += term / theta; }
            if (b) bb += ww[w1] * w[w1] * ihz;
        } else if (b) { bb += w[w1] * ihz; }
        if (k > 0) {
            term = ww[w0] * ihz2;
            double nphi = sdf[c - sz];
            if (nphi < 0) { center += term; }
            else { theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01); center += term / theta; }
            if (b) bb -= ww[w0] * w[w0] * ihz;
        } else if (b) { bb -= w[w0] * ihz; }
        if (b) {
            // moving-boundary contribution (:242-255); boundaryVelocity == 0 when bnd_* is null
            double bc = (1.0 - uw[u1]) * (bu ? bu[u1] : 0.0) * ihx - (1.0 - uw[u0]) * (bu ? bu[u0] : 0.0) * ihx +
                        (1.0 - vw[v1]) * (bv ? bv[v1] : 0.0) * ihy - (1.0 - vw[v0]) * (bv ? bv[v0] : 0.0) * ihy +
                        (1.0 - ww[w1]) * (bw ? bw[w1] : 0.0) * ihz - (1.0 - ww[w0]) * (bw ? bw[w0] : 0.0) * ihz;
            bb += bc;
        }
        if (center < DBL_EPSILON) { center = 1.0; bb = 0.0; }
    } else {
        center = 1.0;
    }
    Acenter[c] = center; Acr[c] = right; Acu[c] = up; Acf[c] = front;
    auto bits = [](float wgt) -> unsigned { return wgt == 0.f ? 0u : (wgt == 1.0f ? 1u : 2u); };
    Acode[c] = (uint8_t)(bits(right) | (bits(up) << 2) | (bits(front) << 4));
    if (b) b[c] = bb;
}

// applyPressureGradient (:715-765): thread (i,j,k) owns faces u(i+1,j,k), v(i,j+1,k), w(i,j,k+1).
__host__ __device__ __forceinline__ void apply_gradient_cell(int nx, int ny, int nz, double hx, double hy, double hz,
                                                             const float* __restrict__ sdf, const float* __restrict__ uw,
                                                             const float* __restrict__ vw, const float* __restrict__ ww,
                                                             const double* __restrict__ x, const double* u_in, const double* v_in,
                                                             const double* w_in, double* u_out, double* v_out, double* w_out, int i,
                                                             int j, int k) {
    const double ihx = 1.0 / hx, ihy = 1.0 / hy, ihz = 1.0 / hz;
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    size_t c = (size_t)k * sz + (size_t)j * sy + i;
    double cphi = sdf[c];
    double xc = x[c];
    if (i + 1 < nx) {
        size_t f = ((size_t)k * ny + j) * (nx + 1) + i + 1;
        double nphi = sdf[c + 1];
        if (uw[f] > 0.0 && (cphi < 0 || nphi < 0)) {
            double theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01);
            u_out[f] = u_in[f] + ihx / theta * (x[c + 1] - xc);
        }
    }
    if (j + 1 < ny) {
        size_t f = ((size_t)k * (ny + 1) + j + 1) * nx + i;
        double nphi = sdf[c + sy];
        if (vw[f] > 0.0 && (cphi < 0 || nphi < 0)) {
            double theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01);
            v_out[f] = v_in[f] + ihy / theta * (x[c + sy] - xc);
        }
    }
    if (k + 1 < nz) {
        size_t f = c + sz;
        double nphi = sdf[c + sz];
        if (ww[f] > 0.0 && (cphi < 0 || nphi < 0)) {
            double theta = fmax(jc_frac_inside_sdf(cphi, nphi), 0.01);
            w_out[f] = w_in[f] + ihz / theta * (x[c + sz] - xc);
        }
    }
}

