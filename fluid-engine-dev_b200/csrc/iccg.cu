// iccg.cu — FdmIccgSolver3 on the device: incomplete-Cholesky preconditioned CG on the 7-point FDM system
// (src/jet/fdm_iccg_solver3.cpp:14-70,152-188; include/jet/detail/cg-inl.h:18-94).  This is the linear solver a
// default-constructed GridFractionalSinglePhasePressureSolver3 carries (…pressure_solver3.cpp:450: 100 iterations, 1e-6) and
// what every hybrid_liquid_sim example runs.
//
// The preconditioner's three recurrences are lexicographic in the reference (inherently serial as written).  A cell depends on
// its lower neighbours only, so the cells of a hyperplane i + j + k = s are independent: the device sweeps 8^3 TILES in tile-
// hyperplane order (one launch per tile hyperplane, one CTA per tile; inside a tile the 22 cell hyperplanes run one after the
// other between __syncthreads) — 94 launches per substitution at 256^3 instead of 766 cell hyperplanes, and exactly the
// reference's arithmetic per cell (bit-exact d, y, x for the same input; tests/test_hostcheck_linear.py, tests/test_gpu_linear.py).
#include "fdm.cuh"
#include "iccg_ops.cuh"

#define IC_T 8

// op: 0 = build d, 1 = forward (out = y), 2 = backward (out = x).  Tile hyperplane S: tiles with ti + tj + tk == S.
// The tile of `out` and the one-cell layer of neighbour values it depends on (below it for build / forward, above it for
// backward: finished by earlier launches) live in shared memory, the cell's couplings, right-hand side and diagonal in registers,
// all fetched up front: a cell hyperplane costs a shared-memory round trip instead of a global one (round 2, GPU call 21).
#define IC_H (IC_T + 1)
template <bool P, int OP>
__global__ void __launch_bounds__(IC_T* IC_T* IC_T)
    k_ic_tiles(Mat A, const double* __restrict__ in, const double* __restrict__ dinv, double* out, int nx, int ny, int nz, int S) {
    jc_pdl_prologue();
    __shared__ double vs[IC_H * IC_H * IC_H];
    const int ti = blockIdx.x, tj = blockIdx.y, tk = S - ti - tj;
    const int ntz = (nz + IC_T - 1) / IC_T;
    if (tk < 0 || tk >= ntz) return;
    const int i0 = ti * IC_T, j0 = tj * IC_T, k0 = tk * IC_T;
    constexpr int LO = OP == 2 ? 0 : 1;   // local coordinate of the tile's first cell inside the (IC_T + 1)^3 block
    for (int e = threadIdx.x; e < IC_H * IC_H * IC_H; e += IC_T * IC_T * IC_T) {
        const int lx = e % IC_H - LO, ly = (e / IC_H) % IC_H - LO, lz = e / (IC_H * IC_H) - LO;
        const bool halo = OP == 2 ? (lx == IC_T || ly == IC_T || lz == IC_T) : (lx < 0 || ly < 0 || lz < 0);
        const int gi = i0 + lx, gj = j0 + ly, gk = k0 + lz;
        const bool in_level = gi >= 0 && gi < nx && gj >= 0 && gj < ny && gk >= 0 && gk < nz;
        vs[e] = (halo && in_level) ? out[((size_t)gk * ny + gj) * nx + gi] : 0.0;   // the tile's own cells are written before read
    }
    const int cx = threadIdx.x % IC_T, cy = (threadIdx.x / IC_T) % IC_T, cz = threadIdx.x / (IC_T * IC_T);
    const int i = i0 + cx, j = j0 + cy, k = k0 + cz;
    const bool live = i < nx && j < ny && k < nz;
    const int mine = OP == 2 ? (3 * IC_T - 3) - (cx + cy + cz) : cx + cy + cz;   // backward: highest local hyperplane first
    const size_t c = ((size_t)k * ny + j) * nx + i;
    const int o = ((cz + LO) * IC_H + (cy + LO)) * IC_H + (cx + LO);
    constexpr int DIR = OP == 2 ? 1 : -1;
    IcCoefs q;
    double rhs = 0.0, dg = 0.0;
    if (live) {
        q = OP == 2 ? ic_upper_coefs<P>(A, i, j, k, nx, ny, nz) : ic_lower_coefs<P>(A, i, j, k, nx, ny);
        rhs = OP == 0 ? A.center[c] : in[c];
        if (OP != 0) dg = dinv[c];
    }
    __syncthreads();
    for (int step = 0; step <= 3 * IC_T - 3; ++step) {
        if (live && mine == step) {
            const double vx = vs[o + DIR], vy = vs[o + DIR * IC_H], vz = vs[o + DIR * IC_H * IC_H];
            vs[o] = OP == 0 ? ic_build_update(q, rhs, vx, vy, vz) : ic_sub_update(q, rhs, vx, vy, vz, dg);
        }
        __syncthreads();
    }
    if (live) out[c] = vs[o];
}

template <bool P, int OP>
static int ic_sweep(jetcu_ctx* ctx, Dim3i d, Mat A, const double* in, const double* dinv, double* out) {
    const int ntx = (d.nx + IC_T - 1) / IC_T, nty = (d.ny + IC_T - 1) / IC_T, ntz = (d.nz + IC_T - 1) / IC_T;
    const int last = ntx + nty + ntz - 3;
    dim3 grid(ntx, nty);
    for (int s = 0; s <= last; ++s) {
        const int S = OP == 2 ? last - s : s;
        JC_LAUNCH_PDL(ctx, s > 0, (k_ic_tiles<P, OP>), grid, IC_T * IC_T * IC_T, 0, A, in, dinv, out, d.nx, d.ny, d.nz, S);
    }
    return JETCU_OK;
}

int iccg_build(jetcu_ctx* ctx, Dim3i d, Mat A, double* dinv) {
    return A.packed ? ic_sweep<true, 0>(ctx, d, A, nullptr, nullptr, dinv) : ic_sweep<false, 0>(ctx, d, A, nullptr, nullptr, dinv);
}

// Preconditioner::solve: y = forward(b), x = backward(y)
int iccg_precondition(jetcu_ctx* ctx, Dim3i d, Mat A, const double* dinv, const double* b, double* y, double* x) {
    if (A.packed) {
        JC_TRY((ic_sweep<true, 1>(ctx, d, A, b, dinv, y)));
        return ic_sweep<true, 2>(ctx, d, A, y, dinv, x);
    }
    JC_TRY((ic_sweep<false, 1>(ctx, d, A, b, dinv, y)));
    return ic_sweep<false, 2>(ctx, d, A, y, dinv, x);
}

static int iccg_read_scalar(jetcu_ctx* ctx, const double* dptr, double* out) {
    JC_TRY(jc_d2h(ctx, ctx->h_scalars, dptr, 1));
    JC_TRY(jc_sync(ctx));
    *out = ctx->h_scalars[0];
    return JETCU_OK;
}

// FdmIccgSolver3::solve (:152-188) = pcg<FdmBlas3, Preconditioner> (cg-inl.h:18-94): x = 0; r = b - Ax; d = M^-1 r; sigma = r.d;
// loop { q = Ad; alpha = sigma / d.q; x += alpha d; r -= alpha q (recomputed when iter % 50 == 0 or after sigma rose);
// s = M^-1 r; sigma' = r.s; d = s + (sigma'/sigma) d } while sigma > tol^2 and iter < maxIter.  Single GPU only.
int iccg_solve_device(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double* x, uint32_t max_iterations, double tolerance,
                      jetcu_mgpcg_result* result) {
    if (jc_comm_world(ctx) > 1)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "ICCG is not slab-distributed: use MGPCG on a multi-GPU communicator");
    const size_t n = d.n();
    const ZRange z{0, d.nz};
    double *r = nullptr, *dd = nullptr, *q = nullptr, *s = nullptr, *dinv = nullptr, *y = nullptr;
    int rc = JETCU_OK;
#define STEP(e) if (rc == JETCU_OK) rc = (e)
    STEP(jc_alloc(ctx, &r, n)); STEP(jc_alloc(ctx, &dd, n)); STEP(jc_alloc(ctx, &q, n)); STEP(jc_alloc(ctx, &s, n));
    STEP(jc_alloc(ctx, &dinv, n)); STEP(jc_alloc(ctx, &y, n));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->stream);
    STEP(jc_zero(ctx, x, n)); STEP(jc_zero(ctx, dd, n)); STEP(jc_zero(ctx, q, n)); STEP(jc_zero(ctx, s, n));
    STEP(jc_zero(ctx, dinv, n)); STEP(jc_zero(ctx, y, n));
    STEP(iccg_build(ctx, d, A, dinv));
    STEP(fdm_residual(ctx, d, z, A, x, b, r));
    STEP(iccg_precondition(ctx, d, A, dinv, r, y, dd));
    double sigmaNew = 0.0;
    STEP(fdm_dot_planes(ctx, d, z, false, r, dd, ctx->d_scalars));
    STEP(iccg_read_scalar(ctx, ctx->d_scalars, &sigmaNew));
    unsigned iter = 0;
    bool trigger = false;
    while (rc == JETCU_OK && sigmaNew > tolerance * tolerance && iter < max_iterations) {
        STEP(fdm_mvm(ctx, d, z, A, dd, q));
        double dq = 0.0;
        STEP(fdm_dot_planes(ctx, d, z, false, dd, q, ctx->d_scalars));
        STEP(iccg_read_scalar(ctx, ctx->d_scalars, &dq));
        const double alpha = sigmaNew / dq;
        const bool recompute = trigger || (iter % 50 == 0 && iter > 0);
        STEP(fdm_pcg_update(ctx, n, alpha, dd, q, x, r, !recompute));
        if (recompute) { STEP(fdm_residual(ctx, d, z, A, x, b, r)); trigger = false; }
        STEP(iccg_precondition(ctx, d, A, dinv, r, y, s));
        const double sigmaOld = sigmaNew;
        STEP(fdm_dot_planes(ctx, d, z, false, r, s, ctx->d_scalars));
        STEP(iccg_read_scalar(ctx, ctx->d_scalars, &sigmaNew));
        if (sigmaNew > sigmaOld) trigger = true;
        const double beta = sigmaNew / sigmaOld;
        STEP(fdm_axpy(ctx, n, beta, dd, s, dd));
        ++iter;
    }
    cudaEventRecord(e1, ctx->stream);
    if (rc == JETCU_OK && result) {
        result->iterations = iter;
        result->last_residual = sqrt(fabs(sigmaNew));
        result->returned = (result->last_residual <= tolerance || iter < max_iterations) ? 1 : 0;
        STEP(fdm_residual(ctx, d, z, A, x, b, r));
        STEP(fdm_dot_planes(ctx, d, z, false, r, r, ctx->d_scalars));
        STEP(fdm_dot_planes(ctx, d, z, false, b, b, ctx->d_scalars + 1));
        STEP(jc_d2h(ctx, ctx->h_scalars, ctx->d_scalars, 2));
        STEP(jc_sync(ctx));
        if (rc == JETCU_OK) {
            result->true_residual_l2 = sqrt(ctx->h_scalars[0]);
            result->b_l2 = sqrt(ctx->h_scalars[1]);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            result->solve_ms = ms;
        }
    } else if (rc == JETCU_OK) {
        rc = jc_sync(ctx);
    }
#undef STEP
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    jc_free(ctx, r); jc_free(ctx, dd); jc_free(ctx, q); jc_free(ctx, s); jc_free(ctx, dinv); jc_free(ctx, y);
    return rc;
}
