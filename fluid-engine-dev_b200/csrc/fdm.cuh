// fdm.cuh — device-side FDM linear algebra (FdmBlas3 / red-black SOR / MG transfer operators / MGPCG driver).
#pragma once
#include "common.cuh"

#define JC_MAX_LEVELS 16

// Plane range [k0, k1) of a level that one launch covers (k is the GLOBAL plane index: arrays are full-size on every
// rank and globally indexed, a rank just works on its slab of them).
struct ZRange {
    int k0, k1;
};

// z-slab decomposition of an MG hierarchy over the ctx's communicator: levels [0, dist_levels) are slab-distributed
// (this rank owns planes [z0[l], z1[l])), coarser levels are replicated (z0 = 0, z1 = nz on every rank).
struct SlabPlan {
    int world = 1, rank = 0;
    int dist_levels = 0;
    int z0[JC_MAX_LEVELS] = {}, z1[JC_MAX_LEVELS] = {};
};

struct MgLevels {
    int levels = 0;
    SlabPlan plan;
    Dim3i dims[JC_MAX_LEVELS];
    Mat A[JC_MAX_LEVELS];
    double* x[JC_MAX_LEVELS] = {};
    double* xalt[JC_MAX_LEVELS] = {};  // ping-pong partner of x (the fused red+black sweep is out of place)
    double* b[JC_MAX_LEVELS] = {};
    double* buf[JC_MAX_LEVELS] = {};
    // allocations owned by this object (the Mat views may also point at caller-owned memory)
    double* own_vec[JC_MAX_LEVELS][4] = {};
    // PCG's search direction d and preconditioned residual s: kept between solves so that their addresses stay put
    double *pcg_d = nullptr, *pcg_s = nullptr;
    // One V-cycle = ~150 small launches (+ ~50 NCCL groups on slabs): captured once per (rhs, x) pair into a CUDA
    // graph and replayed — the solve is launch-bound on the coarse levels otherwise (and on 8 slabs everywhere).
    struct VGraph {
        cudaGraphExec_t exec = nullptr;
        const double* b = nullptr;
        double* x = nullptr;
        uint64_t key = 0;          // hash of everything else the capture baked in (matrices, sweeps, plan)
        uint64_t launches = 0, messages = 0, bytes = 0, stamp = 0;
    } vgraph[2];
    uint64_t vgraph_clock = 0;
};

int fdm_aos_to_soa(jetcu_ctx* ctx, const jetcu_matrix_row* d_aos, MatSoA A, size_t n);
int fdm_mat_to_aos(jetcu_ctx* ctx, Mat A, jetcu_matrix_row* d_aos, size_t n);
int fdm_alloc_mat(jetcu_ctx* ctx, MatSoA* A, size_t n);
void fdm_free_mat(jetcu_ctx* ctx, MatSoA* A);

int fdm_mvm(jetcu_ctx* ctx, Dim3i d, Mat A, const double* v, double* out);
int fdm_mvm(jetcu_ctx* ctx, Dim3i d, ZRange z, Mat A, const double* v, double* out);
int fdm_residual(jetcu_ctx* ctx, Dim3i d, Mat A, const double* x, const double* b, double* out);
int fdm_residual(jetcu_ctx* ctx, Dim3i d, ZRange z, Mat A, const double* x, const double* b, double* out);
// `iterations` red-black SOR sweeps.  On return *x holds the result; *x and *xalt may have been exchanged
// (large grids use an out-of-place fused red+black kernel and ping-pong between the two buffers).
int fdm_relax_rb(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double sor, unsigned iterations, double** x,
                 double** xalt, int level = -1);
// slab form: planes `z` only; `halo` = refresh the planes around `z` from the neighbouring ranks before each colour pass
int fdm_relax_rb(jetcu_ctx* ctx, Dim3i d, ZRange z, bool halo, Mat A, const double* b, double sor, unsigned iterations,
                 double** x, double** xalt, int level = -1);
// `iterations` sweeps of the natural (lexicographic) Gauss-Seidel ordering, in place; single GPU (fdm.cu)
int fdm_relax_lex(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double sor, unsigned iterations, double* x, int level = -1);
int fdm_relax_rhs_halo(int slab_planes, size_t plane_cells);  // planes of b the slab smoother reads outside the slab
int fdm_restrict(jetcu_ctx* ctx, Dim3i fine, const double* finer, double* coarser);
int fdm_restrict(jetcu_ctx* ctx, Dim3i fine, ZRange coarse_planes, const double* finer, double* coarser);
int fdm_correct(jetcu_ctx* ctx, Dim3i fine, const double* coarser, double* finer);
int fdm_correct(jetcu_ctx* ctx, Dim3i fine, ZRange fine_planes, const double* coarser, double* finer);
// slab-invariant dot product over the planes `z` (+ the other ranks' planes when dist); result on the device
int fdm_dot_planes(jetcu_ctx* ctx, Dim3i d, ZRange z, bool dist, const double* a, const double* b, double* d_result);
// result written to d_result[0] (device); deterministic two-stage tree reduction.
int fdm_dot(jetcu_ctx* ctx, size_t n, const double* a, const double* b, double* d_result);
int fdm_axpy(jetcu_ctx* ctx, size_t n, double a, const double* x, const double* y, double* out);
// x += alpha*d ; r -= alpha*q   (the two PCG updates that share alpha, one pass)
int fdm_pcg_update(jetcu_ctx* ctx, size_t n, double alpha, const double* d, const double* q, double* x, double* r,
                   bool update_r);

int mg_alloc_vectors(jetcu_ctx* ctx, MgLevels* mg);
// fills mg->plan from the ctx's communicator (single rank: every level whole)
int mg_make_plan(jetcu_ctx* ctx, MgLevels* mg);
void mg_free(jetcu_ctx* ctx, MgLevels* mg);
int mg_vcycle(jetcu_ctx* ctx, MgLevels* mg, const jetcu_mgpcg_params* p, int level);
// Solves A[0] x = b with MG-preconditioned CG exactly as FdmMgpcgSolver3 does. b, x: device, level-0 size.
int mgpcg_solve_device(jetcu_ctx* ctx, MgLevels* mg, const double* b, double* x, const jetcu_mgpcg_params* p,
                       jetcu_mgpcg_result* result);

// FdmIccgSolver3 (iccg.cu): incomplete-Cholesky preconditioned CG on one level, tile-hyperplane substitutions
int iccg_build(jetcu_ctx* ctx, Dim3i d, Mat A, double* dinv);
int iccg_precondition(jetcu_ctx* ctx, Dim3i d, Mat A, const double* dinv, const double* b, double* y, double* x);
int iccg_solve_device(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double* x, uint32_t max_iterations, double tolerance,
                      jetcu_mgpcg_result* result);
