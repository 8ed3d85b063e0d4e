// diffusion.cu — GridBackwardEulerDiffusionSolver3 on the device (src/jet/grid_backward_euler_diffusion_solver3.cpp): markers
// from the collider / fluid SDFs, the (I - c L) system with Dirichlet or Neumann boundaries, FdmIccgSolver3 (iccg.cu; the
// class default is ICCG(100, kEpsilonD)), for a cell-centred scalar grid or the three arrays of a face-centred grid.
// Callers: GridFluidSolver3::computeViscosity (grid_fluid_solver3.cpp:231-241) of the hybrid and the smoke solver,
// GridSmokeSolver3::computeDiffusion (grid_smoke_solver3.cpp:91-117), and the stand-alone C-ABI entry points.  SURVEY §8f-4.
#include "diffusion.cuh"
#include "diffusion_ops.cuh"

__global__ void __launch_bounds__(256) k_diff_markers(GridD g, DiffArray a, const double* __restrict__ csdf,
                                                      const double* __restrict__ fsdf, char* __restrict__ mk) {
    const int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= a.sx || j >= a.sy) return;
    mk[((size_t)k * a.sy + j) * a.sx + i] = diff_marker(g, a, csdf, fsdf, i, j, k);
}

__global__ void __launch_bounds__(256) k_diff_system(DiffArray a, const char* __restrict__ mk, const double* __restrict__ f,
                                                     double cx, double cy, double cz, int dirichlet, MatSoA A,
                                                     double* __restrict__ b) {
    const int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= a.sx || j >= a.sy) return;
    const size_t c = ((size_t)k * a.sy + j) * a.sx + i;
    double ce, ri, up, fr;
    diff_matrix_row(mk, a.sx, a.sy, a.sz, i, j, k, cx, cy, cz, dirichlet != 0, &ce, &ri, &up, &fr);
    A.center[c] = ce; A.right[c] = ri; A.up[c] = up; A.front[c] = fr;
    b[c] = diff_rhs(mk, f, a.sx, a.sy, a.sz, i, j, k, cx, cy, cz, dirichlet != 0);
}

// One array of GridBackwardEulerDiffusionSolver3::solve; every pointer is device memory, dest may alias nothing else.
int diffusion_solve_array(jetcu_ctx* ctx, const GridD& g, int kind, const double* f, double coefficient, double dt,
                          int boundary_type, const double* csdf, const double* fsdf, uint32_t max_iterations, double tolerance,
                          double* dest, char* markers_out, jetcu_mgpcg_result* result) {
    if (boundary_type != 0 && boundary_type != 1)
        return jetcu_fail(ctx, JETCU_ERR_INVALID_ARG, "diffusion boundary type %d (0 = Dirichlet, 1 = Neumann)", boundary_type);
    const DiffArray a = diff_array(g, kind);
    const size_t n = (size_t)a.sx * a.sy * a.sz;
    // Vector3D c = timeIntervalInSeconds * diffusionCoefficient / (h * h), :34,61,116
    const double tc = dt * coefficient;
    const double cx = tc / (g.hx * g.hx), cy = tc / (g.hy * g.hy), cz = tc / (g.hz * g.hz);
    char* mk = nullptr;
    double *b = nullptr, *x = nullptr;
    MatSoA A;
    int rc = JETCU_OK;
#define STEP(e) if (rc == JETCU_OK) rc = (e)
    STEP(jc_alloc(ctx, &mk, n)); STEP(jc_alloc(ctx, &b, n)); STEP(jc_alloc(ctx, &x, n)); STEP(fdm_alloc_mat(ctx, &A, n));
    if (rc == JETCU_OK) {
        const dim3 grid((a.sx + 63) / 64, (a.sy + 3) / 4, a.sz), blk(64, 4);
        rc = [&]() -> int {
            JC_LAUNCH(ctx, k_diff_markers, grid, blk, 0, g, a, csdf, fsdf, mk);
            JC_LAUNCH(ctx, k_diff_system, grid, blk, 0, a, mk, f, cx, cy, cz, boundary_type == 0 ? 1 : 0, A, b);
            return JETCU_OK;
        }();
    }
    STEP(iccg_solve_device(ctx, Dim3i{a.sx, a.sy, a.sz}, mat_of(A), b, x, max_iterations, tolerance, result));
    STEP(jc_d2d(ctx, dest, x, n));
    if (markers_out) STEP(jc_d2d(ctx, markers_out, mk, n));
#undef STEP
    jc_free(ctx, mk); jc_free(ctx, b); jc_free(ctx, x); fdm_free_mat(ctx, &A);
    return rc;
}

// solve(const FaceCenteredGrid3&, ...), :108-160: u, v, w one after the other with their own markers.  out may alias in
// only if the caller passes copies as `in` (GridFluidSolver3::computeViscosity clones the velocity first).
int diffusion_solve_faces(jetcu_ctx* ctx, const GridD& g, const double* u, const double* v, const double* w, double coefficient,
                          double dt, int boundary_type, const double* csdf, const double* fsdf, uint32_t max_iterations,
                          double tolerance, double* ou, double* ov, double* ow, jetcu_mgpcg_result results[3]) {
    jetcu_mgpcg_result r[3] = {};
    JC_TRY(diffusion_solve_array(ctx, g, 1, u, coefficient, dt, boundary_type, csdf, fsdf, max_iterations, tolerance, ou, nullptr, &r[0]));
    JC_TRY(diffusion_solve_array(ctx, g, 2, v, coefficient, dt, boundary_type, csdf, fsdf, max_iterations, tolerance, ov, nullptr, &r[1]));
    JC_TRY(diffusion_solve_array(ctx, g, 3, w, coefficient, dt, boundary_type, csdf, fsdf, max_iterations, tolerance, ow, nullptr, &r[2]));
    if (results) for (int q = 0; q < 3; ++q) results[q] = r[q];
    return JETCU_OK;
}

// ---------------------------------------------------------------------------------------------- C ABI (host buffers)
namespace {
struct DBuf {   // device copy of an optional host array
    jetcu_ctx* ctx;
    double* p = nullptr;
    explicit DBuf(jetcu_ctx* c) : ctx(c) {}
    ~DBuf() { jc_free(ctx, p); }
    int up(const double* h, size_t n) {
        JC_TRY(jc_alloc(ctx, &p, n));
        return h ? jc_h2d(ctx, p, h, n) : JETCU_OK;
    }
};
int diff_check_grid(jetcu_ctx* ctx, const jetcu_grid* grid, GridD* g) {
    JC_ARG(ctx, grid != nullptr);
    JC_ARG(ctx, grid->nx > 0 && grid->ny > 0 && grid->nz > 0 && grid->nx < (1u << 15) && grid->ny < (1u << 15) && grid->nz < (1u << 15));
    JC_ARG(ctx, grid->hx > 0.0 && grid->hy > 0.0 && grid->hz > 0.0);
    *g = GridD{(int)grid->nx, (int)grid->ny, (int)grid->nz, grid->hx, grid->hy, grid->hz, grid->ox, grid->oy, grid->oz,
               jc_pow2_inverse(grid->hx), jc_pow2_inverse(grid->hy), jc_pow2_inverse(grid->hz)};
    return JETCU_OK;
}
}  // namespace

extern "C" {

int jetcu_diffusion_solve_faces(jetcu_ctx* ctx, const jetcu_grid* grid, const double* u, const double* v, const double* w,
                                double coefficient, double dt, int boundary_type, const double* collider_sdf,
                                const double* fluid_sdf, uint32_t max_iterations, double tolerance, double* u_out, double* v_out,
                                double* w_out, jetcu_mgpcg_result results[3]) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");
    GridD g;
    JC_TRY(diff_check_grid(ctx, grid, &g));
    JC_ARG(ctx, u && v && w && u_out && v_out && w_out);
    const size_t nc = (size_t)g.nx * g.ny * g.nz;
    const size_t nf[3] = {(size_t)(g.nx + 1) * g.ny * g.nz, (size_t)g.nx * (g.ny + 1) * g.nz, (size_t)g.nx * g.ny * (g.nz + 1)};
    DBuf du(ctx), dv(ctx), dw(ctx), ou(ctx), ov(ctx), ow(ctx), cs(ctx), fs(ctx);
    JC_TRY(du.up(u, nf[0])); JC_TRY(dv.up(v, nf[1])); JC_TRY(dw.up(w, nf[2]));
    JC_TRY(ou.up(nullptr, nf[0])); JC_TRY(ov.up(nullptr, nf[1])); JC_TRY(ow.up(nullptr, nf[2]));
    if (collider_sdf) JC_TRY(cs.up(collider_sdf, nc));
    if (fluid_sdf) JC_TRY(fs.up(fluid_sdf, nc));
    JC_TRY(diffusion_solve_faces(ctx, g, du.p, dv.p, dw.p, coefficient, dt, boundary_type, cs.p, fs.p, max_iterations, tolerance,
                                 ou.p, ov.p, ow.p, results));
    JC_TRY(jc_d2h(ctx, u_out, ou.p, nf[0])); JC_TRY(jc_d2h(ctx, v_out, ov.p, nf[1])); JC_TRY(jc_d2h(ctx, w_out, ow.p, nf[2]));
    return jc_sync(ctx);
}

int jetcu_diffusion_solve_cells(jetcu_ctx* ctx, const jetcu_grid* grid, const double* source, double coefficient, double dt,
                                int boundary_type, const double* collider_sdf, const double* fluid_sdf, uint32_t max_iterations,
                                double tolerance, double* dest, char* markers, jetcu_mgpcg_result* result) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");
    GridD g;
    JC_TRY(diff_check_grid(ctx, grid, &g));
    JC_ARG(ctx, source && dest);
    const size_t nc = (size_t)g.nx * g.ny * g.nz;
    DBuf src(ctx), dst(ctx), cs(ctx), fs(ctx);
    JC_TRY(src.up(source, nc)); JC_TRY(dst.up(nullptr, nc));
    if (collider_sdf) JC_TRY(cs.up(collider_sdf, nc));
    if (fluid_sdf) JC_TRY(fs.up(fluid_sdf, nc));
    char* dmk = nullptr;
    if (markers) JC_TRY(jc_alloc(ctx, &dmk, nc));
    jetcu_mgpcg_result r{};
    int rc = diffusion_solve_array(ctx, g, 0, src.p, coefficient, dt, boundary_type, cs.p, fs.p, max_iterations, tolerance, dst.p, dmk, &r);
    if (rc == JETCU_OK) rc = jc_d2h(ctx, dest, dst.p, nc);
    if (rc == JETCU_OK && markers) rc = jc_d2h(ctx, markers, dmk, nc);
    if (rc == JETCU_OK) rc = jc_sync(ctx);
    jc_free(ctx, dmk);
    if (rc == JETCU_OK && result) *result = r;
    return rc;
}

}  // extern "C"
