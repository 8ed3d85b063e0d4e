// fdm.cuh — device-side FDM linear algebra (FdmBlas3 / red-black SOR / MG transfer operators / MGPCG driver).
#pragma once
#include "common.cuh"

#define JC_MAX_LEVELS 16

struct MgLevels {
    int levels = 0;
    Dim3i dims[JC_MAX_LEVELS];
    Mat A[JC_MAX_LEVELS];
    double* x[JC_MAX_LEVELS] = {};
    double* xalt[JC_MAX_LEVELS] = {};  // ping-pong partner of x (the fused red+black sweep is out of place)
    double* b[JC_MAX_LEVELS] = {};
    double* buf[JC_MAX_LEVELS] = {};
    // allocations owned by this object (the Mat views may also point at caller-owned memory)
    double* own_vec[JC_MAX_LEVELS][4] = {};
};

int fdm_aos_to_soa(jetcu_ctx* ctx, const jetcu_matrix_row* d_aos, MatSoA A, size_t n);
int fdm_mat_to_aos(jetcu_ctx* ctx, Mat A, jetcu_matrix_row* d_aos, size_t n);
int fdm_alloc_mat(jetcu_ctx* ctx, MatSoA* A, size_t n);
void fdm_free_mat(jetcu_ctx* ctx, MatSoA* A);

int fdm_mvm(jetcu_ctx* ctx, Dim3i d, Mat A, const double* v, double* out);
int fdm_residual(jetcu_ctx* ctx, Dim3i d, Mat A, const double* x, const double* b, double* out);
// `iterations` red-black SOR sweeps.  On return *x holds the result; *x and *xalt may have been exchanged
// (large grids use an out-of-place fused red+black kernel and ping-pong between the two buffers).
int fdm_relax_rb(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double sor, unsigned iterations, double** x,
                 double** xalt, int level = -1);
int fdm_restrict(jetcu_ctx* ctx, Dim3i fine, const double* finer, double* coarser);
int fdm_residual_restrict(jetcu_ctx* ctx, Dim3i fine, Mat A, const double* x, const double* b, double* buf,
                          double* coarser);
int fdm_correct(jetcu_ctx* ctx, Dim3i fine, const double* coarser, double* finer);
// result written to d_result[0] (device); deterministic two-stage tree reduction.
int fdm_dot(jetcu_ctx* ctx, size_t n, const double* a, const double* b, double* d_result);
int fdm_axpy(jetcu_ctx* ctx, size_t n, double a, const double* x, const double* y, double* out);
// x += alpha*d ; r -= alpha*q   (the two PCG updates that share alpha, one pass)
int fdm_pcg_update(jetcu_ctx* ctx, size_t n, double alpha, const double* d, const double* q, double* x, double* r,
                   bool update_r);

int mg_alloc_vectors(jetcu_ctx* ctx, MgLevels* mg);
void mg_free(jetcu_ctx* ctx, MgLevels* mg);
int mg_vcycle(jetcu_ctx* ctx, MgLevels* mg, const jetcu_mgpcg_params* p, int level);
// Solves A[0] x = b with MG-preconditioned CG exactly as FdmMgpcgSolver3 does. b, x: device, level-0 size.
int mgpcg_solve_device(jetcu_ctx* ctx, MgLevels* mg, const double* b, double* x, const jetcu_mgpcg_params* p,
                       jetcu_mgpcg_result* result);
