// pressure.cu — GridFractionalSinglePhasePressureSolver3 on the device.
// Reference: src/jet/grid_fractional_single_phase_pressure_solver3.cpp (line ranges cited per kernel).
// No reductions anywhere in the system build => sdf/weights/A/b are bit-identical to the CPU path.
#include "pressure.cuh"
#include "pressure_ops.cuh"

#define PBX 64
#define PBY 4
static inline dim3 pgrid(int nx, int ny, int nz) { return dim3((nx + PBX - 1) / PBX, (ny + PBY - 1) / PBY, nz); }

// Kernel bodies: pressure_ops.cuh (__host__ __device__, one cell / face each).
__global__ void __launch_bounds__(PBX* PBY) k_fluid_sdf0(GridP g, const double* __restrict__ fluid,
                                                         float* __restrict__ out) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    fluid_sdf0_cell(g, fluid, out, i, j, k);
}

template <int AXIS>
__global__ void __launch_bounds__(PBX* PBY) k_face_weights0(GridP g, const double* __restrict__ collider,
                                                            float* __restrict__ out) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1), sz = g.nz + (AXIS == 2);
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= sx || j >= sy || k >= sz) return;
    face_weight0_face<AXIS>(g, collider, out, i, j, k);
}

__global__ void __launch_bounds__(PBX* PBY) k_restrict_f32(const float* __restrict__ finer, int fx, int fy, int fz,
                                                           float* __restrict__ coarser, int cx, int cy, int cz) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= cx || j >= cy || k >= cz) return;
    restrict_f32_cell(finer, fx, fy, fz, coarser, cx, cy, cz, i, j, k);
}

__global__ void __launch_bounds__(PBX* PBY)
    k_build_system(int nx, int ny, int nz, double hx, double hy, double hz, const float* __restrict__ sdf,
                   const float* __restrict__ uw, const float* __restrict__ vw, const float* __restrict__ ww,
                   const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                   const double* __restrict__ bu, const double* __restrict__ bv, const double* __restrict__ bw,
                   double* __restrict__ Acenter, float* __restrict__ Acr, float* __restrict__ Acu, float* __restrict__ Acf,
                   uint8_t* __restrict__ Acode, double* __restrict__ b) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= nx || j >= ny) return;
    build_system_cell(nx, ny, nz, hx, hy, hz, sdf, uw, vw, ww, u, v, w, bu, bv, bw, Acenter, Acr, Acu, Acf, Acode, b, i, j, k);
}

__global__ void __launch_bounds__(PBX* PBY)
    k_apply_gradient(int nx, int ny, int nz, double hx, double hy, double hz, const float* __restrict__ sdf,
                     const float* __restrict__ uw, const float* __restrict__ vw, const float* __restrict__ ww,
                     const double* __restrict__ x, const double* u_in, const double* v_in, const double* w_in,
                     double* u_out, double* v_out, double* w_out) {
    int i = blockIdx.x * PBX + threadIdx.x, j = blockIdx.y * PBY + threadIdx.y, k = blockIdx.z;
    if (i >= nx || j >= ny) return;
    apply_gradient_cell(nx, ny, nz, hx, hy, hz, sdf, uw, vw, ww, x, u_in, v_in, w_in, u_out, v_out, w_out, i, j, k);
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
    // FdmIccgSolver3 (Jet's default) works on the single level-0 system — buildSingleSystem is the same function in both paths
    // (…pressure_solver3.cpp:651-713) — and so does the device.  useCompressedLinearSystem(true) (JETCU_SOLVER_COMPRESSED)
    // asks the reference to drop the air rows (buildCompressedSystem, :269-444); they are identity rows with b = 0 that no
    // liquid row couples to, so the liquid-cell solution is the same and the device keeps the uncompressed layout.
    if (p->linear_solver & JETCU_SOLVER_ICCG)
        JC_TRY(iccg_solve_device(ctx, ps->dims[0], ps->mg.A[0], ps->b0, ps->x0, p->max_iterations, p->tolerance, result));
    else
        JC_TRY(mgpcg_solve_device(ctx, &ps->mg, ps->b0, ps->x0, p, result));
    JC_LAUNCH(ctx, k_apply_gradient, pgrid(g.nx, g.ny, g.nz), blk, 0, g.nx, g.ny, g.nz, g.hx, g.hy, g.hz, ps->sdf[0],
              ps->uw[0], ps->vw[0], ps->ww[0], ps->x0, u_in, v_in, w_in, u_out, v_out, w_out);
    return JETCU_OK;
}
