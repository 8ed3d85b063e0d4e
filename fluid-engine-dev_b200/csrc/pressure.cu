// pressure.cu — GridFractionalSinglePhasePressureSolver3 on the device.
// Reference: src/jet/grid_fractional_single_phase_pressure_solver3.cpp (line ranges cited per kernel).
// No reductions anywhere in the system build => sdf/weights/A/b are bit-identical to the CPU path.
#include "pressure.cuh"
#include "sampling.cuh"

#define PBX 64
#define PBY 4
static inline dim3 pgrid(int nx, int ny, int nz) { return dim3((nx + PBX - 1) / PBX, (ny + PBY - 1) / PBY, nz); }

struct GridP {
    int nx, ny, nz;
    double hx, hy, hz, ox, oy, oz;
};

// _fluidSdf[0](i,j,k) = float(fluidSdf.sample(cellCenterPosition(i,j,k)))            (:546-549)
// cellCenter = o + h*(i+0.5) (src/jet/grid3.cpp:32-38); cell-centred data origin = o + 0.5*h
// (src/jet/cell_centered_scalar_grid3.cpp:60-62).  fluid == nullptr => ConstantScalarField3(-kMaxD).
__global__ void __launch_bounds__(PBX* PBY) k_fluid_sdf0(GridP g, const double* __restrict__ fluid,
                                                         float* __restrict__ out) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    double val;
    if (fluid) {
        GridView v{fluid, g.nx, g.ny, g.nz, g.hx, g.hy, g.hz, g.ox + 0.5 * g.hx, g.oy + 0.5 * g.hy, g.oz + 0.5 * g.hz};
        val = jc_sample(v, g.ox + g.hx * (i + 0.5), g.oy + g.hy * (j + 0.5), g.oz + g.hz * (k + 0.5));
    } else {
        val = -DBL_MAX;
    }
    out[((size_t)k * g.ny + j) * g.nx + i] = (float)val;
}

__device__ __forceinline__ float weight_from_frac(double frac) {
    // clamp(1 - frac, 0, 1), values in (0, kMinWeight) raised to kMinWeight = 0.01   (:563-571)
    double weight = jc_clamp(1.0 - frac, 0.0, 1.0);
    if (weight < 0.01 && weight > 0.0) weight = 0.01;
    return (float)weight;
}

// _{u,v,w}Weights[0]: open fraction of each face from the collider SDF at its 4 corners (:551-617).
// AXIS = 0/1/2 for u/v/w.  Face position = faceOrigin + h*(i,j,k) (src/jet/face_centered_grid3.cpp:216-238,442-444).
template <int AXIS>
__global__ void __launch_bounds__(PBX* PBY) k_face_weights0(GridP g, const double* __restrict__ collider,
                                                            float* __restrict__ out) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1), sz = g.nz + (AXIS == 2);
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= sx || j >= sy || k >= sz) return;
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
__global__ void __launch_bounds__(PBX* PBY) k_restrict_f32(const float* __restrict__ finer, int fx, int fy, int fz,
                                                           float* __restrict__ coarser, int cx, int cy, int cz) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= cx || j >= cy || k >= cz) return;
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
__global__ void __launch_bounds__(PBX* PBY)
    k_build_system(int nx, int ny, int nz, double hx, double hy, double hz, const float* __restrict__ sdf,
                   const float* __restrict__ uw, const float* __restrict__ vw, const float* __restrict__ ww,
                   const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                   const double* __restrict__ bu, const double* __restrict__ bv, const double* __restrict__ bw,
                   double* __restrict__ Acenter, float* __restrict__ Acr, float* __restrict__ Acu, float* __restrict__ Acf,
                   uint8_t* __restrict__ Acode, double* __restrict__ b) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= nx || j >= ny) return;
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
__global__ void __launch_bounds__(PBX* PBY)
    k_apply_gradient(int nx, int ny, int nz, double hx, double hy, double hz, const float* __restrict__ sdf,
                     const float* __restrict__ uw, const float* __restrict__ vw, const float* __restrict__ ww,
                     const double* __restrict__ x, const double* u_in, const double* v_in, const double* w_in,
                     double* u_out, double* v_out, double* w_out) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= nx || j >= ny) return;
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

void pressure_free(jetcu_ctx* ctx, PressureSystem* ps) {
    for (int l = 0; l < ps->levels; ++l) {
        jc_free(ctx, ps->sdf[l]); jc_free(ctx, ps->uw[l]); jc_free(ctx, ps->vw[l]); jc_free(ctx, ps->ww[l]);
        jc_free(ctx, ps->center[l]); jc_free(ctx, ps->codes[l]); jc_free(ctx, ps->code8[l]);
    }
    mg_free(ctx, &ps->mg);
    jc_free(ctx, ps->b0); jc_free(ctx, ps->x0);
    ps->levels = 0;
}

int pressure_prepare(jetcu_ctx* ctx, PressureSystem* ps, const jetcu_grid* g, uint32_t max_levels) {
    if (ps->levels > 0 && ps->grid.nx == g->nx && ps->grid.ny == g->ny && ps->grid.nz == g->nz &&
        ps->built_max_levels == max_levels) {
        ps->grid = *g;
        return JETCU_OK;
    }
    pressure_free(ctx, ps);
    ps->grid = *g;
    ps->built_max_levels = max_levels;
    int levels = (int)jetcu_mg_num_levels(g->nx, g->ny, g->nz, max_levels);
    if (levels > JC_MAX_LEVELS) return jetcu_fail(ctx, JETCU_ERR_INVALID_ARG, "too many MG levels (%d)", levels);
    ps->levels = levels;
    ps->mg.levels = levels;
    for (int l = 0; l < levels; ++l) {
        Dim3i d{(int)(g->nx >> l), (int)(g->ny >> l), (int)(g->nz >> l)};
        ps->dims[l] = d;
        ps->mg.dims[l] = d;
        JC_TRY(jc_alloc(ctx, &ps->sdf[l], d.n()));
        JC_TRY(jc_alloc(ctx, &ps->uw[l], (size_t)(d.nx + 1) * d.ny * d.nz));
        JC_TRY(jc_alloc(ctx, &ps->vw[l], (size_t)d.nx * (d.ny + 1) * d.nz));
        JC_TRY(jc_alloc(ctx, &ps->ww[l], (size_t)d.nx * d.ny * (d.nz + 1)));
        JC_TRY(jc_alloc(ctx, &ps->center[l], d.n()));
        JC_TRY(jc_alloc(ctx, &ps->codes[l], 3 * d.n()));
        JC_TRY(jc_alloc(ctx, &ps->code8[l], d.n() + 16));
    }
    JC_TRY(mg_alloc_vectors(ctx, &ps->mg));
    JC_TRY(jc_alloc(ctx, &ps->b0, ps->dims[0].n()));
    JC_TRY(jc_alloc(ctx, &ps->x0, ps->dims[0].n()));
    return JETCU_OK;
}

int pressure_solve_device(jetcu_ctx* ctx, PressureSystem* ps, const double* u_in, const double* v_in,
                          const double* w_in, const double* collider_sdf, const double* bnd_u, const double* bnd_v,
                          const double* bnd_w, const double* fluid_sdf, const jetcu_mgpcg_params* p, double* u_out,
                          double* v_out, double* w_out, jetcu_mgpcg_result* result) {
    const jetcu_grid& gg = ps->grid;
    GridP g{(int)gg.nx, (int)gg.ny, (int)gg.nz, gg.hx, gg.hy, gg.hz, gg.ox, gg.oy, gg.oz};
    dim3 blk(PBX, PBY);
    // buildWeights (:520-636)
    JC_LAUNCH(ctx, k_fluid_sdf0, pgrid(g.nx, g.ny, g.nz), blk, 0, g, fluid_sdf, ps->sdf[0]);
    JC_LAUNCH(ctx, k_face_weights0<0>, pgrid(g.nx + 1, g.ny, g.nz), blk, 0, g, collider_sdf, ps->uw[0]);
    JC_LAUNCH(ctx, k_face_weights0<1>, pgrid(g.nx, g.ny + 1, g.nz), blk, 0, g, collider_sdf, ps->vw[0]);
    JC_LAUNCH(ctx, k_face_weights0<2>, pgrid(g.nx, g.ny, g.nz + 1), blk, 0, g, collider_sdf, ps->ww[0]);
    for (int l = 1; l < ps->levels; ++l) {
        Dim3i f = ps->dims[l - 1], c = ps->dims[l];
        JC_LAUNCH(ctx, k_restrict_f32, pgrid(c.nx, c.ny, c.nz), blk, 0, ps->sdf[l - 1], f.nx, f.ny, f.nz, ps->sdf[l], c.nx, c.ny, c.nz);
        JC_LAUNCH(ctx, k_restrict_f32, pgrid(c.nx + 1, c.ny, c.nz), blk, 0, ps->uw[l - 1], f.nx + 1, f.ny, f.nz, ps->uw[l], c.nx + 1, c.ny, c.nz);
        JC_LAUNCH(ctx, k_restrict_f32, pgrid(c.nx, c.ny + 1, c.nz), blk, 0, ps->vw[l - 1], f.nx, f.ny + 1, f.nz, ps->vw[l], c.nx, c.ny + 1, c.nz);
        JC_LAUNCH(ctx, k_restrict_f32, pgrid(c.nx, c.ny, c.nz + 1), blk, 0, ps->ww[l - 1], f.nx, f.ny, f.nz + 1, ps->ww[l], c.nx, c.ny, c.nz + 1);
    }
    // buildSystem (:651-713): level l uses spacing h * 2^l
    double hx = g.hx, hy = g.hy, hz = g.hz;
    for (int l = 0; l < ps->levels; ++l) {
        Dim3i d = ps->dims[l];
        JC_LAUNCH(ctx, k_build_system, pgrid(d.nx, d.ny, d.nz), blk, 0, d.nx, d.ny, d.nz, hx, hy, hz, ps->sdf[l], ps->uw[l],
                  ps->vw[l], ps->ww[l], l == 0 ? u_in : nullptr, l == 0 ? v_in : nullptr, l == 0 ? w_in : nullptr,
                  l == 0 ? bnd_u : nullptr, l == 0 ? bnd_v : nullptr, l == 0 ? bnd_w : nullptr, ps->center[l],
                  ps->codes[l], ps->codes[l] + d.n(), ps->codes[l] + 2 * d.n(), ps->code8[l], l == 0 ? ps->b0 : nullptr);
        Mat m;
        m.packed = true;
        m.center = ps->center[l];
        m.cr = ps->codes[l]; m.cu = ps->codes[l] + d.n(); m.cf = ps->codes[l] + 2 * d.n();
        m.code = ps->code8[l];
        m.stride_y = (size_t)d.nx; m.stride_z = (size_t)d.nx * d.ny;
        // term = weight * invHSqr with invHSqr = invH * invH (…pressure_solver3.cpp:127-128)
        m.sx = (1.0 / hx) * (1.0 / hx); m.sy = (1.0 / hy) * (1.0 / hy); m.sz = (1.0 / hz) * (1.0 / hz);
        {   // k_build_system's accumulation order for a bulk cell: +x, -x, +y, -y, +z, -z, every term = 1.0f * invH^2
            volatile double c = 0.0;
            c += m.sx; c += m.sx; c += m.sy; c += m.sy; c += m.sz; c += m.sz;
            m.c6 = c;
        }
        ps->mg.A[l] = m;
        hx *= 2.0; hy *= 2.0; hz *= 2.0;
    }
    JC_TRY(mgpcg_solve_device(ctx, &ps->mg, ps->b0, ps->x0, p, result));
    JC_LAUNCH(ctx, k_apply_gradient, pgrid(g.nx, g.ny, g.nz), blk, 0, g.nx, g.ny, g.nz, g.hx, g.hy, g.hz, ps->sdf[0],
              ps->uw[0], ps->vw[0], ps->ww[0], ps->x0, u_in, v_in, w_in, u_out, v_out, w_out);
    return JETCU_OK;
}
