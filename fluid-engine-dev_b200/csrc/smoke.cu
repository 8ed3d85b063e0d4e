// smoke.cu — device-resident GridSmokeSolver3 (BASELINE config 5) behind include/jetcu.h.
//
// Mirrors GridFluidSolver3::onAdvanceTimeStep (src/jet/grid_fluid_solver3.cpp:179-210) with GridSmokeSolver3's overrides
// (src/jet/grid_smoke_solver3.cpp): boundary condition -> buoyancy -> pressure projection (the same
// GridFractionalSinglePhasePressureSolver3 + MGPCG path as the liquid solver, on an all-liquid domain) -> cubic
// semi-Lagrangian advection of density, temperature and velocity -> decay.  No emitter, no collider, zero diffusion /
// viscosity coefficients (the solver's defaults).  Kernel bodies live in smoke_ops.cuh (__host__ __device__; checked on the
// CPU by tests/hostcheck); this file holds the launch wrappers, the ambient-temperature reduction and the host side.
// [Written after round 1's GPU minutes ran out: tests/firstrun/zz_smoke.py is its first GPU run.]
#include <math.h>
#include <string.h>

#include "pressure.cuh"
#include "smoke_ops.cuh"
#include "diffusion.cuh"

struct jetcu_smoke {
    jetcu_ctx* ctx = nullptr;
    jetcu_smoke_params P{};
    int nx = 0, ny = 0, nz = 0;
    size_t nu = 0, nv = 0, nw = 0, nc = 0;
    double *u = nullptr, *v = nullptr, *w = nullptr;        // velocity
    double *u0 = nullptr, *v0 = nullptr, *w0 = nullptr;     // the clone the advection reads
    double *den = nullptr, *temp = nullptr, *scratch = nullptr;
    // rigid collider (GridFluidSolver3::setCollider), col.kind == 0: none — the same pieces as in the hybrid solver
    ColliderD col;
    double* col_sdf = nullptr;                                    // GridFractionalBoundaryConditionSolver3::_colliderSdf
    double *bnd_u = nullptr, *bnd_v = nullptr, *bnd_w = nullptr;  // collider velocity at the face positions (pressure rhs)
    double *tmp_u = nullptr, *tmp_v = nullptr, *tmp_w = nullptr;  // uTemp / vTemp / wTemp of constrainVelocity
    char *cm_u = nullptr, *cm_v = nullptr, *cm_w = nullptr;       // its markers; cm_u doubles as the scalar extrapolation marker
    char *valid0 = nullptr, *valid1 = nullptr;                    // ping-pong validity maps of extrapolateToRegion
    PressureSystem ps;
    jetcu_mgpcg_result last{};
    cudaEvent_t ev[5] = {};
    double ms_forces = 0, ms_pressure = 0, ms_advection = 0, ms_decay = 0;
    // GridFluidSolver3::viscosityCoefficient, GridSmokeSolver3::smokeDiffusionCoefficient / temperatureDiffusionCoefficient
    double viscosity = 0.0, smoke_diffusion = 0.0, temperature_diffusion = 0.0;
};

static GridD smoke_grid(const jetcu_smoke* h) {
    const jetcu_grid& g = h->P.grid;
    return GridD{h->nx, h->ny, h->nz, g.hx, g.hy, g.hz, g.ox, g.oy, g.oz,
                 jc_pow2_inverse(g.hx), jc_pow2_inverse(g.hy), jc_pow2_inverse(g.hz)};
}

static inline dim3 sgrid(int sx, int sy, int sz) { return dim3((sx + 63) / 64, (sy + 3) / 4, sz); }

__global__ void k_smoke_zero_walls(GridD g, int flag, double* __restrict__ u, double* __restrict__ v, double* __restrict__ w) {
    zero_walls_elem(g, flag, u, v, w, (size_t)blockIdx.x * blockDim.x + threadIdx.x);
}

template <int AXIS>
__global__ void __launch_bounds__(256) k_smoke_buoyancy(GridD g, const double* __restrict__ den, const double* __restrict__ temp,
                                                        double fd, double ft, const double* __restrict__ t_sum, double inv_n,
                                                        double dt, double up_a, double* __restrict__ comp) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx + (AXIS == 0) || j >= g.ny + (AXIS == 1) || k >= g.nz + (AXIS == 2)) return;
    // tAmb = sum / N  (grid_smoke_solver3.cpp:147-151; inv_n is passed as N itself: the reference DIVIDES)
    sm_buoyancy_face<AXIS>(g, den, temp, fd, ft, t_sum[0] / inv_n, dt, up_a, comp, i, j, k);
}

// k0: first plane of this launch (z-slab mode: a rank advects the planes it owns, the fields are all-gathered afterwards)
__global__ void __launch_bounds__(256) k_smoke_advect_cell(GridD g, const double* __restrict__ in, const double* __restrict__ u,
                                                           const double* __restrict__ v, const double* __restrict__ w, double dt,
                                                           double hmin, bool linear, const double* __restrict__ csdf,
                                                           double* __restrict__ out, int k0) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = k0 + blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    out[((size_t)k * g.ny + j) * g.nx + i] = sm_advect_cell(g, in, u, v, w, dt, hmin, i, j, k, linear, csdf);
}

template <int AXIS>
__global__ void __launch_bounds__(256) k_smoke_advect_face(GridD g, const double* __restrict__ u0, const double* __restrict__ v0,
                                                           const double* __restrict__ w0, double dt, double hmin, bool linear,
                                                           const double* __restrict__ csdf, double* __restrict__ out, int k0) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1), sz = g.nz + (AXIS == 2);
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = k0 + blockIdx.z;
    if (i >= sx || j >= sy || k >= sz) return;
    out[((size_t)k * sy + j) * sx + i] = sm_advect_face<AXIS>(g, u0, v0, w0, dt, hmin, i, j, k, linear, csdf);
}

__global__ void __launch_bounds__(256) k_smoke_emit_step(GridD g, ColliderD e, double width, double lo, double hi, double* __restrict__ field) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    sm_emit_step_cell(g, e, width, lo, hi, field, i, j, k);
}

__global__ void k_smoke_scale(double* __restrict__ a, double s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] *= s;
}

// Sum of the temperature field for tAmb.  The reference adds the cells serially (k, j, i order); here a fixed two-stage
// tree (592 CTAs, then one) — deterministic, differs from the serial sum by rounding only.
#define SMOKE_SUM_GRID 592
__global__ void __launch_bounds__(256) k_smoke_sum_partial(const double* __restrict__ a, size_t n, double* __restrict__ partials) {
    __shared__ double sm[256];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)SMOKE_SUM_GRID * 256) acc += a[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) k_smoke_sum_final(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double sm[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partials[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// GridFluidSolver3::cfl, src/jet/grid_fluid_solver3.cpp:57-72: max over cells of the SIGNED components of valueAtCellCenter + dt*g,
// starting from 0 (max is exact: any reduction order gives the same bits) — the liquid solver's k_cfl_partial, local to this TU
__global__ void __launch_bounds__(256) k_smoke_cfl_partial(GridD g, const double* __restrict__ u, const double* __restrict__ v,
                                                           const double* __restrict__ w, double dtgx, double dtgy, double dtgz,
                                                           double* __restrict__ partials) {
    __shared__ double sm[256];
    double m = 0.0;
    const size_t nc = (size_t)g.nx * g.ny * g.nz;
    for (size_t c = (size_t)blockIdx.x * 256 + threadIdx.x; c < nc; c += (size_t)SMOKE_SUM_GRID * 256) m = fmax(m, sm_cfl_cell(g, u, v, w, dtgx, dtgy, dtgz, c));
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) k_smoke_max_final(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double sm[256];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) m = fmax(m, partials[i]);
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// Collider pieces: thin launch wrappers around collider_ops.cuh (the bodies hybrid.cu's kernels wrap too; checked on the CPU by
// tests/hostcheck).  Kept local to this TU so that the liquid solver's translation unit is not touched.
__global__ void __launch_bounds__(256) k_smk_collider_sdf(GridD g, ColliderD c, double* __restrict__ sdf) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    sdf[((size_t)k * g.ny + j) * g.nx + i] = collider_sdf_cell(g, c, i, j, k);
}
template <int AXIS>
__global__ void __launch_bounds__(256) k_smk_boundary_velocity(GridD g, ColliderD c, double* __restrict__ bnd) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx + (AXIS == 0) || j >= g.ny + (AXIS == 1) || k >= g.nz + (AXIS == 2)) return;
    boundary_velocity_face<AXIS>(g, c, bnd, i, j, k);
}
template <int AXIS>
__global__ void __launch_bounds__(256) k_smk_constrain_markers(GridD g, ColliderD c, const double* __restrict__ csdf,
                                                               double* __restrict__ data, char* __restrict__ marker) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx + (AXIS == 0) || j >= g.ny + (AXIS == 1) || k >= g.nz + (AXIS == 2)) return;
    constrain_marker_face<AXIS>(g, c, csdf, data, marker, i, j, k);
}
template <int AXIS>
__global__ void __launch_bounds__(256) k_smk_constrain_project(GridD g, ColliderD c, const double* __restrict__ csdf,
                                                               const double* __restrict__ u, const double* __restrict__ v,
                                                               const double* __restrict__ w, double* __restrict__ temp) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx + (AXIS == 0) || j >= g.ny + (AXIS == 1) || k >= g.nz + (AXIS == 2)) return;
    constrain_project_face<AXIS>(g, c, csdf, u, v, w, temp, i, j, k);
}
__global__ void __launch_bounds__(256) k_smk_outside_marker(GridD g, const double* __restrict__ csdf, char* __restrict__ marker) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    marker[((size_t)k * g.ny + j) * g.nx + i] = outside_collider_cell(g, csdf, i, j, k);
}
__global__ void __launch_bounds__(256) k_smk_extrapolate_sweep(double* data, const char* __restrict__ v0, char* __restrict__ v1, int nx,
                                                               int ny, int nz) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= nx || j >= ny) return;
    extrapolate_cell(data, v0, v1, nx, ny, nz, i, j, k);
}

#define SMK_ENTER(h)                                                      \
    if (!(h)) return JETCU_ERR_INVALID_ARG;                               \
    jetcu_ctx* ctx = (h)->ctx;                                            \
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");

// extrapolateToRegion (include/jet/detail/array_utils-inl.h:191-239), `depth` sweeps from the validity map `marker`
static int smoke_extrapolate(jetcu_smoke* h, double* data, const char* marker, int sx, int sy, int sz, unsigned depth) {
    jetcu_ctx* ctx = h->ctx;
    JC_TRY(jc_d2d(ctx, h->valid0, marker, (size_t)sx * sy * sz));
    char *v0 = h->valid0, *v1 = h->valid1;
    for (unsigned it = 0; it < depth; ++it) {
        JC_LAUNCH(ctx, k_smk_extrapolate_sweep, sgrid(sx, sy, sz), dim3(64, 4), 0, data, v0, v1, sx, sy, sz);
        char* t = v0; v0 = v1; v1 = t;
    }
    return JETCU_OK;
}

// GridFluidSolver3::applyBoundaryCondition == constrainVelocity: the collider branch (…boundary_condition_solver3.cpp:55-187),
// then the closed walls (:190-232)
static int smoke_apply_bc(jetcu_smoke* h) {
    jetcu_ctx* ctx = h->ctx;
    GridD g = smoke_grid(h);
    if (h->col.kind != 0) {
        const unsigned depth = (unsigned)ceil(h->P.max_cfl);  // grid_fluid_solver3.cpp:346-353
        const dim3 blk(64, 4);
        const int nx = h->nx, ny = h->ny, nz = h->nz;
        JC_LAUNCH(ctx, k_smk_constrain_markers<0>, sgrid(nx + 1, ny, nz), blk, 0, g, h->col, h->col_sdf, h->u, h->cm_u);
        JC_LAUNCH(ctx, k_smk_constrain_markers<1>, sgrid(nx, ny + 1, nz), blk, 0, g, h->col, h->col_sdf, h->v, h->cm_v);
        JC_LAUNCH(ctx, k_smk_constrain_markers<2>, sgrid(nx, ny, nz + 1), blk, 0, g, h->col, h->col_sdf, h->w, h->cm_w);
        JC_TRY(smoke_extrapolate(h, h->u, h->cm_u, nx + 1, ny, nz, depth));
        JC_TRY(smoke_extrapolate(h, h->v, h->cm_v, nx, ny + 1, nz, depth));
        JC_TRY(smoke_extrapolate(h, h->w, h->cm_w, nx, ny, nz + 1, depth));
        JC_LAUNCH(ctx, k_smk_constrain_project<0>, sgrid(nx + 1, ny, nz), blk, 0, g, h->col, h->col_sdf, h->u, h->v, h->w, h->tmp_u);
        JC_LAUNCH(ctx, k_smk_constrain_project<1>, sgrid(nx, ny + 1, nz), blk, 0, g, h->col, h->col_sdf, h->u, h->v, h->w, h->tmp_v);
        JC_LAUNCH(ctx, k_smk_constrain_project<2>, sgrid(nx, ny, nz + 1), blk, 0, g, h->col, h->col_sdf, h->u, h->v, h->w, h->tmp_w);
        JC_TRY(jc_d2d(ctx, h->u, h->tmp_u, h->nu)); JC_TRY(jc_d2d(ctx, h->v, h->tmp_v, h->nv)); JC_TRY(jc_d2d(ctx, h->w, h->tmp_w, h->nw));
    }
    size_t m = (size_t)g.ny * g.nz;
    if ((size_t)g.nx * g.nz > m) m = (size_t)g.nx * g.nz;
    if ((size_t)g.nx * g.ny > m) m = (size_t)g.nx * g.ny;
    JC_LAUNCH(ctx, k_smoke_zero_walls, jc_blocks(m, 256), 256, 0, g, h->P.closed_domain_boundary_flag, h->u, h->v, h->w);
    return JETCU_OK;
}

extern "C" {

void jetcu_smoke_default_params(jetcu_smoke_params* p) {
    memset(p, 0, sizeof(*p));
    p->grid.nx = p->grid.ny = p->grid.nz = 1;
    p->grid.hx = p->grid.hy = p->grid.hz = 1.0;
    p->closed_domain_boundary_flag = JETCU_DIR_ALL;
    p->gravity[1] = -9.8;
    p->max_cfl = 5.0;
    p->buoyancy_density_factor = -0.000625; p->buoyancy_temperature_factor = 5.0;   // include/jet/grid_smoke_solver3.h:172-175
    p->density_decay = 0.001; p->temperature_decay = 0.001;
    p->mgpcg.max_iterations = 100; p->mgpcg.max_levels = 5; p->mgpcg.n_restriction = 5; p->mgpcg.n_correction = 5;
    p->mgpcg.n_coarsest = 20; p->mgpcg.n_final = 20; p->mgpcg.tolerance = 1e-9; p->mgpcg.sor_factor = 1.5;
    p->mgpcg.use_red_black = 1;
}

int jetcu_smoke_destroy(jetcu_smoke* h) {
    if (!h) return JETCU_OK;
    jetcu_ctx* ctx = h->ctx;
    cudaSetDevice(ctx->device);
    jc_free(ctx, h->u); jc_free(ctx, h->v); jc_free(ctx, h->w); jc_free(ctx, h->u0); jc_free(ctx, h->v0); jc_free(ctx, h->w0);
    jc_free(ctx, h->den); jc_free(ctx, h->temp); jc_free(ctx, h->scratch);
    jc_free(ctx, h->col_sdf); jc_free(ctx, h->bnd_u); jc_free(ctx, h->bnd_v); jc_free(ctx, h->bnd_w);
    jc_free(ctx, h->tmp_u); jc_free(ctx, h->tmp_v); jc_free(ctx, h->tmp_w);
    jc_free(ctx, h->cm_u); jc_free(ctx, h->cm_v); jc_free(ctx, h->cm_w); jc_free(ctx, h->valid0); jc_free(ctx, h->valid1);
    pressure_free(ctx, &h->ps);
    cudaStreamSynchronize(ctx->stream);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    delete h;
    return JETCU_OK;
}

int jetcu_smoke_create(jetcu_ctx* ctx, const jetcu_smoke_params* p, jetcu_smoke** out) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");
    JC_ARG(ctx, p && out);
    JC_ARG(ctx, p->grid.nx > 0 && p->grid.ny > 0 && p->grid.nz > 0 && p->grid.nx < 32768 && p->grid.ny < 32768 && p->grid.nz < 32768);
    JC_ARG(ctx, p->grid.hx > 0 && p->grid.hy > 0 && p->grid.hz > 0);
    if (!p->mgpcg.use_red_black && jc_comm_world(ctx) > 1)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "lexicographic Gauss-Seidel (useRedBlackOrdering = false) is not slab-distributed: "
                          "use useRedBlackOrdering = true on a multi-GPU communicator");
    JC_ARG(ctx, p->advection == JETCU_ADVECT_CUBIC || p->advection == JETCU_ADVECT_LINEAR);
    jetcu_smoke* h = new jetcu_smoke();
    h->ctx = ctx;
    h->P = *p;
    h->nx = (int)p->grid.nx; h->ny = (int)p->grid.ny; h->nz = (int)p->grid.nz;
    h->nu = (size_t)(h->nx + 1) * h->ny * h->nz; h->nv = (size_t)h->nx * (h->ny + 1) * h->nz;
    h->nw = (size_t)h->nx * h->ny * (h->nz + 1); h->nc = (size_t)h->nx * h->ny * h->nz;
    int rc = JETCU_OK;
#define STEP(x) if (rc == JETCU_OK) rc = (x)
    STEP(jc_alloc(ctx, &h->u, h->nu)); STEP(jc_alloc(ctx, &h->v, h->nv)); STEP(jc_alloc(ctx, &h->w, h->nw));
    STEP(jc_alloc(ctx, &h->u0, h->nu)); STEP(jc_alloc(ctx, &h->v0, h->nv)); STEP(jc_alloc(ctx, &h->w0, h->nw));
    STEP(jc_alloc(ctx, &h->den, h->nc)); STEP(jc_alloc(ctx, &h->temp, h->nc)); STEP(jc_alloc(ctx, &h->scratch, h->nc));
    STEP(jc_zero(ctx, h->u, h->nu)); STEP(jc_zero(ctx, h->v, h->nv)); STEP(jc_zero(ctx, h->w, h->nw));
    STEP(jc_zero(ctx, h->den, h->nc)); STEP(jc_zero(ctx, h->temp, h->nc));
    STEP(pressure_prepare(ctx, &h->ps, &p->grid, p->mgpcg.max_levels));
    STEP(jc_zero(ctx, h->ps.x0, h->nc));
    STEP(jc_sync(ctx));
#undef STEP
    if (rc != JETCU_OK) { jetcu_smoke_destroy(h); return rc; }
    *out = h;
    return JETCU_OK;
}

int jetcu_smoke_set_fields(jetcu_smoke* h, const double* den, const double* temp, const double* u, const double* v, const double* w) {
    SMK_ENTER(h);
    if (den) JC_TRY(jc_h2d(ctx, h->den, den, h->nc));
    if (temp) JC_TRY(jc_h2d(ctx, h->temp, temp, h->nc));
    if (u) JC_TRY(jc_h2d(ctx, h->u, u, h->nu));
    if (v) JC_TRY(jc_h2d(ctx, h->v, v, h->nv));
    if (w) JC_TRY(jc_h2d(ctx, h->w, w, h->nw));
    return jc_sync(ctx);
}

int jetcu_smoke_get_fields(jetcu_smoke* h, double* den, double* temp, double* u, double* v, double* w, double* pr) {
    SMK_ENTER(h);
    if (den) JC_TRY(jc_d2h(ctx, den, h->den, h->nc));
    if (temp) JC_TRY(jc_d2h(ctx, temp, h->temp, h->nc));
    if (u) JC_TRY(jc_d2h(ctx, u, h->u, h->nu));
    if (v) JC_TRY(jc_d2h(ctx, v, h->v, h->nv));
    if (w) JC_TRY(jc_d2h(ctx, w, h->w, h->nw));
    if (pr) JC_TRY(jc_d2h(ctx, pr, h->ps.x0, h->nc));
    return jc_sync(ctx);
}

// GridFluidSolver3::cfl / numberOfSubTimeSteps (src/jet/grid_fluid_solver3.cpp:57-72, 212-217) on the device-resident velocity
int jetcu_smoke_cfl(jetcu_smoke* h, double dt, double* cfl) {
    SMK_ENTER(h);
    JC_ARG(ctx, cfl);
    GridD g = smoke_grid(h);
    const double* gr = h->P.gravity;
    JC_LAUNCH(ctx, k_smoke_cfl_partial, SMOKE_SUM_GRID, 256, 0, g, h->u, h->v, h->w, dt * gr[0], dt * gr[1], dt * gr[2], ctx->d_partials);
    JC_LAUNCH(ctx, k_smoke_max_final, 1, 256, 0, ctx->d_partials, SMOKE_SUM_GRID, ctx->d_scalars + 17);
    JC_TRY(jc_d2h(ctx, ctx->h_scalars, ctx->d_scalars + 17, 1));
    JC_TRY(jc_sync(ctx));
    *cfl = ctx->h_scalars[0] * dt / fmin(fmin(g.hx, g.hy), g.hz);
    return JETCU_OK;
}
int jetcu_smoke_num_sub_steps(jetcu_smoke* h, double dt, uint32_t* n) {
    SMK_ENTER(h);
    JC_ARG(ctx, n);
    double c;
    JC_TRY(jetcu_smoke_cfl(h, dt, &c));
    *n = (uint32_t)fmax(ceil(c / h->P.max_cfl), 1.0);
    return JETCU_OK;
}

// GridFluidSolver3::setCollider(RigidBodyCollider3 over analytic surfaces) on the smoke solver
int jetcu_smoke_set_collider_shapes(jetcu_smoke* h, const jetcu_collider_shape* shapes, uint32_t n, int as_set,
                                    const double linear_velocity[3], const double angular_velocity[3], double friction) {
    SMK_ENTER(h);
    JC_ARG(ctx, n <= JC_COLLIDER_MAX_SHAPES && (n == 0 || shapes != nullptr) && (n <= 1 || as_set));
    for (uint32_t i = 0; i < n; ++i) {
        JC_ARG(ctx, shapes[i].kind >= JETCU_COLLIDER_BOX && shapes[i].kind <= JETCU_COLLIDER_PLANE);
        if (shapes[i].kind == JETCU_COLLIDER_SPHERE || shapes[i].kind == JETCU_COLLIDER_CYLINDER) JC_ARG(ctx, shapes[i].b[0] > 0.0);
    }
    h->col = ColliderD();
    if (n == 0) return JETCU_OK;
    const ColliderD c = cd_make_collider(shapes, n, as_set, linear_velocity, angular_velocity, friction);
    if (!h->col_sdf) {
        JC_TRY(jc_alloc(ctx, &h->col_sdf, h->nc));
        JC_TRY(jc_alloc(ctx, &h->bnd_u, h->nu)); JC_TRY(jc_alloc(ctx, &h->bnd_v, h->nv)); JC_TRY(jc_alloc(ctx, &h->bnd_w, h->nw));
        JC_TRY(jc_alloc(ctx, &h->tmp_u, h->nu)); JC_TRY(jc_alloc(ctx, &h->tmp_v, h->nv)); JC_TRY(jc_alloc(ctx, &h->tmp_w, h->nw));
        JC_TRY(jc_alloc(ctx, &h->cm_u, h->nu)); JC_TRY(jc_alloc(ctx, &h->cm_v, h->nv)); JC_TRY(jc_alloc(ctx, &h->cm_w, h->nw));
        size_t m = h->nu > h->nv ? h->nu : h->nv;
        if (h->nw > m) m = h->nw;
        JC_TRY(jc_alloc(ctx, &h->valid0, m)); JC_TRY(jc_alloc(ctx, &h->valid1, m));
    }
    GridD g = smoke_grid(h);
    const dim3 blk(64, 4);
    // re-rasterised by the reference at every beginAdvanceTimeStep (grid_fluid_solver3.cpp:448-453): a pure function of the
    // (static) shape, so once here
    JC_LAUNCH(ctx, k_smk_collider_sdf, sgrid(h->nx, h->ny, h->nz), blk, 0, g, c, h->col_sdf);
    JC_LAUNCH(ctx, k_smk_boundary_velocity<0>, sgrid(h->nx + 1, h->ny, h->nz), blk, 0, g, c, h->bnd_u);
    JC_LAUNCH(ctx, k_smk_boundary_velocity<1>, sgrid(h->nx, h->ny + 1, h->nz), blk, 0, g, c, h->bnd_v);
    JC_LAUNCH(ctx, k_smk_boundary_velocity<2>, sgrid(h->nx, h->ny, h->nz + 1), blk, 0, g, c, h->bnd_w);
    JC_TRY(jc_sync(ctx));
    h->col = c;
    return JETCU_OK;
}

// VolumeGridEmitter3 (step-function targets on smoke density and temperature) fired once on the device
int jetcu_smoke_emit_step(jetcu_smoke* h, const jetcu_collider_shape* shapes, uint32_t n, int as_set, double density_min,
                          double density_max, double temperature_min, double temperature_max) {
    SMK_ENTER(h);
    JC_ARG(ctx, shapes && n >= 1 && n <= JC_COLLIDER_MAX_SHAPES && (n == 1 || as_set));
    for (uint32_t i = 0; i < n; ++i) {
        JC_ARG(ctx, shapes[i].kind >= JETCU_COLLIDER_BOX && shapes[i].kind <= JETCU_COLLIDER_PLANE);
        if (shapes[i].kind == JETCU_COLLIDER_SPHERE || shapes[i].kind == JETCU_COLLIDER_CYLINDER) JC_ARG(ctx, shapes[i].b[0] > 0.0);
    }
    const ColliderD e = cd_make_collider(shapes, n, as_set, nullptr, nullptr, 0.0);
    GridD g = smoke_grid(h);
    const double width = cd_min(cd_min(g.hx, g.hy), g.hz);   // scalarGridTarget->gridSpacing().min()
    const dim3 blk(64, 4);
    JC_LAUNCH(ctx, k_smoke_emit_step, sgrid(h->nx, h->ny, h->nz), blk, 0, g, e, width, density_min, density_max, h->den);
    JC_LAUNCH(ctx, k_smoke_emit_step, sgrid(h->nx, h->ny, h->nz), blk, 0, g, e, width, temperature_min, temperature_max, h->temp);
    return jc_sync(ctx);
}

// GridSmokeSolver3::computeExternalForces -> computeBuoyancyForce (src/jet/grid_smoke_solver3.cpp:91-93,131-187)
int jetcu_smoke_compute_external_forces(jetcu_smoke* h, double dt) {
    SMK_ENTER(h);
    GridD g = smoke_grid(h);
    const double* gr = h->P.gravity;
    double up[3] = {0, 1, 0};
    if (gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2] > DBL_EPSILON) {
        const double l = sqrt(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
        up[0] = -(gr[0] / l); up[1] = -(gr[1] / l); up[2] = -(gr[2] / l);
    }
    const double fd = h->P.buoyancy_density_factor, ft = h->P.buoyancy_temperature_factor;
    if (fabs(fd) > DBL_EPSILON || fabs(ft) > DBL_EPSILON) {
        JC_LAUNCH(ctx, k_smoke_sum_partial, SMOKE_SUM_GRID, 256, 0, h->temp, h->nc, ctx->d_partials);
        JC_LAUNCH(ctx, k_smoke_sum_final, 1, 256, 0, ctx->d_partials, SMOKE_SUM_GRID, ctx->d_scalars + 16);
        const double ncells = (double)h->nc;
        const dim3 blk(64, 4);
        if (fabs(up[0]) > DBL_EPSILON)
            JC_LAUNCH(ctx, k_smoke_buoyancy<0>, sgrid(h->nx + 1, h->ny, h->nz), blk, 0, g, h->den, h->temp, fd, ft, ctx->d_scalars + 16,
                      ncells, dt, up[0], h->u);
        if (fabs(up[1]) > DBL_EPSILON)
            JC_LAUNCH(ctx, k_smoke_buoyancy<1>, sgrid(h->nx, h->ny + 1, h->nz), blk, 0, g, h->den, h->temp, fd, ft, ctx->d_scalars + 16,
                      ncells, dt, up[1], h->v);
        if (fabs(up[2]) > DBL_EPSILON)
            JC_LAUNCH(ctx, k_smoke_buoyancy<2>, sgrid(h->nx, h->ny, h->nz + 1), blk, 0, g, h->den, h->temp, fd, ft, ctx->d_scalars + 16,
                      ncells, dt, up[2], h->w);
        JC_TRY(smoke_apply_bc(h));
    }
    return JETCU_OK;
}

// GridFluidSolver3::computePressure (grid_fluid_solver3.cpp:243-253): fluidSdf = -kMaxD (:313-315), collider grid = kMaxD
int jetcu_smoke_compute_pressure(jetcu_smoke* h, double dt, jetcu_mgpcg_result* result) {
    SMK_ENTER(h);
    (void)dt;
    const bool col = h->col.kind != 0;
    JC_TRY(pressure_solve_device(ctx, &h->ps, h->u, h->v, h->w, col ? h->col_sdf : nullptr, col ? h->bnd_u : nullptr,
                                 col ? h->bnd_v : nullptr, col ? h->bnd_w : nullptr, nullptr, &h->P.mgpcg, h->u, h->v, h->w, &h->last));
    if (result) *result = h->last;
    return smoke_apply_bc(h);
}

// z-slab mode: the planes [z0, z1) this rank advects.  Storage stays replicated (full-size, globally indexed arrays on
// every rank, like the liquid solver's grids): the semi-Lagrangian back-trace reads up to maxCfl + 2 planes away, which a
// replicated input serves without a halo protocol; only the OUTPUT planes are exchanged (one all-gather per field).
static void smoke_slab(const jetcu_smoke* h, int* z0, int* z1, int* world, int* rank) {
    *world = jc_comm_world(h->ctx); *rank = jc_comm_rank(h->ctx);
    if (*world > 1 && h->nz % *world == 0) { const int T = h->nz / *world; *z0 = *rank * T; *z1 = *z0 + T; }
    else { *z0 = 0; *z1 = h->nz; *world = 1; *rank = 0; }
}

static int smoke_gather_planes(jetcu_smoke* h, int world, int z0, int z1, void* field, size_t plane_bytes, bool extra_top_plane) {
    if (world <= 1) return JETCU_OK;
    JC_TRY(comm_allgather_inplace(h->ctx, field, (size_t)(z1 - z0) * plane_bytes));
    if (extra_top_plane) JC_TRY(comm_bcast(h->ctx, static_cast<char*>(field) + (size_t)h->nz * plane_bytes, plane_bytes, world - 1));
    return JETCU_OK;
}

// GridFluidSolver3::computeAdvection (grid_fluid_solver3.cpp:255-311) with CubicSemiLagrangian3
int jetcu_smoke_compute_advection(jetcu_smoke* h, double dt) {
    SMK_ENTER(h);
    GridD g = smoke_grid(h);
    const double hmin = fmin(fmin(g.hx, g.hy), g.hz);
    const bool linear = h->P.advection == JETCU_ADVECT_LINEAR;
    const double* csdf = h->col.kind != 0 ? h->col_sdf : nullptr;   // boundarySdf = *colliderSdf() (kMaxD grid without collider)
    const dim3 blk(64, 4);
    int z0, z1, world, rank;
    smoke_slab(h, &z0, &z1, &world, &rank);
    const int nk = z1 - z0, nkw = nk + (rank == world - 1 ? 1 : 0);   // the last rank also owns the top w plane
    const size_t pc = (size_t)h->nx * h->ny * 8, pu = (size_t)(h->nx + 1) * h->ny * 8, pv = (size_t)h->nx * (h->ny + 1) * 8;
    // advectable scalars in registration order (density, temperature): input = clone, flow = the current velocity
    double* fields[2] = {h->den, h->temp};
    for (double* f : fields) {
        JC_TRY(jc_d2d(ctx, h->scratch, f, h->nc));
        JC_LAUNCH(ctx, k_smoke_advect_cell, sgrid(h->nx, h->ny, nk), blk, 0, g, h->scratch, h->u, h->v, h->w, dt, hmin, linear, csdf, f, z0);
        JC_TRY(smoke_gather_planes(h, world, z0, z1, f, pc, false));
        if (csdf) {  // GridFluidSolver3::extrapolateIntoCollider(ScalarGrid3*), grid_fluid_solver3.cpp:355-369
            JC_LAUNCH(ctx, k_smk_outside_marker, sgrid(h->nx, h->ny, h->nz), blk, 0, g, csdf, h->cm_u);
            JC_TRY(smoke_extrapolate(h, f, h->cm_u, h->nx, h->ny, h->nz, (unsigned)ceil(h->P.max_cfl)));
        }
    }
    // velocity: input = flow = the clone
    JC_TRY(jc_d2d(ctx, h->u0, h->u, h->nu)); JC_TRY(jc_d2d(ctx, h->v0, h->v, h->nv)); JC_TRY(jc_d2d(ctx, h->w0, h->w, h->nw));
    JC_LAUNCH(ctx, k_smoke_advect_face<0>, sgrid(h->nx + 1, h->ny, nk), blk, 0, g, h->u0, h->v0, h->w0, dt, hmin, linear, csdf, h->u, z0);
    JC_LAUNCH(ctx, k_smoke_advect_face<1>, sgrid(h->nx, h->ny + 1, nk), blk, 0, g, h->u0, h->v0, h->w0, dt, hmin, linear, csdf, h->v, z0);
    JC_LAUNCH(ctx, k_smoke_advect_face<2>, sgrid(h->nx, h->ny, nkw), blk, 0, g, h->u0, h->v0, h->w0, dt, hmin, linear, csdf, h->w, z0);
    JC_TRY(smoke_gather_planes(h, world, z0, z1, h->u, pu, false));
    JC_TRY(smoke_gather_planes(h, world, z0, z1, h->v, pv, false));
    JC_TRY(smoke_gather_planes(h, world, z0, z1, h->w, pc, true));
    return smoke_apply_bc(h);
}

// setViscosityCoefficient / setSmokeDiffusionCoefficient / setTemperatureDiffusionCoefficient
int jetcu_smoke_set_diffusion(jetcu_smoke* h, double viscosity, double smoke_diffusion, double temperature_diffusion) {
    if (!h) return JETCU_ERR_INVALID_ARG;
    if (!(viscosity >= 0.0 && smoke_diffusion >= 0.0 && temperature_diffusion >= 0.0))
        return jetcu_fail(h->ctx, JETCU_ERR_INVALID_ARG, "negative diffusion coefficient");
    h->viscosity = viscosity; h->smoke_diffusion = smoke_diffusion; h->temperature_diffusion = temperature_diffusion;
    return JETCU_OK;
}

// GridFluidSolver3::computeViscosity (grid_fluid_solver3.cpp:231-241): fluidSdf() of a grid fluid solver is
// ConstantScalarField3(-kMaxD), so every sample outside the collider is fluid.  Single GPU (ICCG).
int jetcu_smoke_compute_viscosity(jetcu_smoke* h, double dt) {
    SMK_ENTER(h);
    if (!(h->viscosity > DBL_EPSILON)) return JETCU_OK;
    JC_TRY(jc_d2d(ctx, h->u0, h->u, h->nu)); JC_TRY(jc_d2d(ctx, h->v0, h->v, h->nv)); JC_TRY(jc_d2d(ctx, h->w0, h->w, h->nw));
    JC_TRY(diffusion_solve_faces(ctx, smoke_grid(h), h->u0, h->v0, h->w0, h->viscosity, dt, 1, h->col.kind != 0 ? h->col_sdf : nullptr,
                                 nullptr, 100, DBL_EPSILON, h->u, h->v, h->w, nullptr));
    return smoke_apply_bc(h);
}

// one scalar of GridSmokeSolver3::computeDiffusion (grid_smoke_solver3.cpp:91-117): solve on a clone, extrapolateIntoCollider
static int smoke_diffuse_scalar(jetcu_smoke* h, double* field, double coefficient, double dt) {
    jetcu_ctx* ctx = h->ctx;
    const GridD g = smoke_grid(h);
    const double* csdf = h->col.kind != 0 ? h->col_sdf : nullptr;
    JC_TRY(jc_d2d(ctx, h->scratch, field, h->nc));
    JC_TRY(diffusion_solve_array(ctx, g, 0, h->scratch, coefficient, dt, 1, csdf, nullptr, 100, DBL_EPSILON, field, nullptr, nullptr));
    if (csdf) {
        JC_LAUNCH(ctx, k_smk_outside_marker, sgrid(h->nx, h->ny, h->nz), dim3(64, 4), 0, g, csdf, h->cm_u);
        JC_TRY(smoke_extrapolate(h, field, h->cm_u, h->nx, h->ny, h->nz, (unsigned)ceil(h->P.max_cfl)));
    }
    return JETCU_OK;
}

// GridSmokeSolver3::onEndAdvanceTimeStep -> computeDiffusion (:85-128): the scalar diffusion solves where their coefficients are
// > kEpsilonD, then the decays
int jetcu_smoke_end_advance_time_step(jetcu_smoke* h, double dt) {
    SMK_ENTER(h);
    if (h->smoke_diffusion > DBL_EPSILON) JC_TRY(smoke_diffuse_scalar(h, h->den, h->smoke_diffusion, dt));
    if (h->temperature_diffusion > DBL_EPSILON) JC_TRY(smoke_diffuse_scalar(h, h->temp, h->temperature_diffusion, dt));
    JC_LAUNCH(ctx, k_smoke_scale, jc_blocks(h->nc, 256), 256, 0, h->den, 1.0 - h->P.density_decay, h->nc);
    JC_LAUNCH(ctx, k_smoke_scale, jc_blocks(h->nc, 256), 256, 0, h->temp, 1.0 - h->P.temperature_decay, h->nc);
    return JETCU_OK;
}

// GridFluidSolver3::onAdvanceTimeStep (grid_fluid_solver3.cpp:179-210)
int jetcu_smoke_step(jetcu_smoke* h, double dt, jetcu_mgpcg_result* result) {
    SMK_ENTER(h);
    if (!h->ev[0]) for (auto& e : h->ev) JC_CUDA(ctx, cudaEventCreate(&e));
    cudaStream_t st = ctx->stream;
    cudaEventRecord(h->ev[0], st);
    JC_TRY(smoke_apply_bc(h));  // beginAdvanceTimeStep: applyBoundaryCondition (:455-456)
    JC_TRY(jetcu_smoke_compute_external_forces(h, dt));
    JC_TRY(jetcu_smoke_compute_viscosity(h, dt));   // no-op unless set_diffusion(viscosity > kEpsilonD)
    cudaEventRecord(h->ev[1], st);
    JC_TRY(jetcu_smoke_compute_pressure(h, dt, result));
    cudaEventRecord(h->ev[2], st);
    JC_TRY(jetcu_smoke_compute_advection(h, dt));
    cudaEventRecord(h->ev[3], st);
    JC_TRY(jetcu_smoke_end_advance_time_step(h, dt));
    cudaEventRecord(h->ev[4], st);
    JC_TRY(jc_sync(ctx));
    float ms[4] = {0, 0, 0, 0};
    for (int q = 0; q < 4; ++q) cudaEventElapsedTime(&ms[q], h->ev[q], h->ev[q + 1]);
    h->ms_forces = ms[0]; h->ms_pressure = ms[1]; h->ms_advection = ms[2]; h->ms_decay = ms[3];
    return JETCU_OK;
}

// device time of the stages of the last jetcu_smoke_step (CUDA events on the library's stream), milliseconds:
// buoyancy + boundary condition, pressure projection, advection (all five fields), decay
int jetcu_smoke_last_step_ms(jetcu_smoke* h, double* forces, double* pressure, double* advection, double* decay) {
    SMK_ENTER(h);
    if (forces) *forces = h->ms_forces;
    if (pressure) *pressure = h->ms_pressure;
    if (advection) *advection = h->ms_advection;
    if (decay) *decay = h->ms_decay;
    return JETCU_OK;
}

}  // extern "C"
