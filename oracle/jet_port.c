/*
 * jet_port.c — TEST INFRASTRUCTURE ("oracle", kind = "port").
 *
 * A plain-C restatement of the arithmetic on Jet's hybrid-liquid hot path (doyubkim/fluid-engine-dev),
 * written from the behaviour of the reference files cited at each function.  It exists only to CHECK the
 * CUDA path: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never
 * links or calls it.  Parity status: PINNED — tests/test_oracle_*.py check every function here against the
 * unmodified reference compiled from /root/reference (oracle/_ref/libjetref.so, see oracle/Makefile) and
 * against the golden vectors under tests/golden/ generated from it (tests/golden/make_golden.py).
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off (no FMA contraction: the reference is built for baseline x86-64).
 */
#include "jet_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef ptrdiff_t ssz;

static int g_threads = 0;

const char* jeto_kind(void) { return "port"; }
void jeto_set_num_threads(int n) {
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_num_threads(n);
#else
    n = 1;
#endif
    g_threads = n;
}
int jeto_num_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

#define IDX(i, j, k, nx, ny) (((size_t)(k) * (ny) + (size_t)(j)) * (nx) + (size_t)(i))

/* ====================================================================== FdmBlas3 */

/* FdmBlas3::mvm — src/jet/fdm_linear_system3.cpp:79-96.  Term order: center, -i, +i, -j, +j, -k, +k. */
int jeto_fdm_mvm(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* v, double* out) {
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < (ssz)nz; ++k)
        for (ssz j = 0; j < (ssz)ny; ++j)
            for (ssz i = 0; i < (ssz)nx; ++i) {
                size_t c = IDX(i, j, k, nx, ny);
                out[c] = A[c].center * v[c] +
                         ((i > 0) ? A[c - 1].right * v[c - 1] : 0.0) +
                         ((i + 1 < (ssz)nx) ? A[c].right * v[c + 1] : 0.0) +
                         ((j > 0) ? A[c - nx].up * v[c - nx] : 0.0) +
                         ((j + 1 < (ssz)ny) ? A[c].up * v[c + nx] : 0.0) +
                         ((k > 0) ? A[c - nx * ny].front * v[c - nx * ny] : 0.0) +
                         ((k + 1 < (ssz)nz) ? A[c].front * v[c + nx * ny] : 0.0);
            }
    return 0;
}

/* FdmBlas3::residual — src/jet/fdm_linear_system3.cpp:98-116. */
int jeto_fdm_residual(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* x, const double* b,
                      double* out) {
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < (ssz)nz; ++k)
        for (ssz j = 0; j < (ssz)ny; ++j)
            for (ssz i = 0; i < (ssz)nx; ++i) {
                size_t c = IDX(i, j, k, nx, ny);
                out[c] = b[c] - A[c].center * x[c] -
                         ((i > 0) ? A[c - 1].right * x[c - 1] : 0.0) -
                         ((i + 1 < (ssz)nx) ? A[c].right * x[c + 1] : 0.0) -
                         ((j > 0) ? A[c - nx].up * x[c - nx] : 0.0) -
                         ((j + 1 < (ssz)ny) ? A[c].up * x[c + nx] : 0.0) -
                         ((k > 0) ? A[c - nx * ny].front * x[c - nx * ny] : 0.0) -
                         ((k + 1 < (ssz)nz) ? A[c].front * x[c + nx * ny] : 0.0);
            }
    return 0;
}

/* FdmBlas3::dot — src/jet/fdm_linear_system3.cpp:49-65: serial k-j-i accumulation. */
int jeto_fdm_dot(uint64_t n, const double* a, const double* b, double* result) {
    double r = 0.0;
    for (uint64_t i = 0; i < n; ++i) r += a[i] * b[i];
    *result = r;
    return 0;
}

/* FdmBlas3::axpy — src/jet/fdm_linear_system3.cpp:67-77. */
int jeto_fdm_axpy(uint64_t n, double a, const double* x, const double* y, double* out) {
#pragma omp parallel for schedule(static)
    for (ssz i = 0; i < (ssz)n; ++i) out[i] = a * x[i] + y[i];
    return 0;
}

/* One colour of FdmGaussSeidelSolver3::relaxRedBlack — src/jet/fdm_gauss_seidel_solver3.cpp:150-221.
 * colour 0 = red = (i+j+k) even. */
static void relax_colour(ssz nx, ssz ny, ssz nz, const jeto_row* A, const double* b, double w, double* x,
                         int colour) {
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j) {
            ssz i = colour == 0 ? (j + k) % 2 : 1 - (j + k) % 2;
            for (; i < nx; i += 2) {
                size_t c = IDX(i, j, k, nx, ny);
                double r = ((i > 0) ? A[c - 1].right * x[c - 1] : 0.0) +
                           ((i + 1 < nx) ? A[c].right * x[c + 1] : 0.0) +
                           ((j > 0) ? A[c - nx].up * x[c - nx] : 0.0) +
                           ((j + 1 < ny) ? A[c].up * x[c + nx] : 0.0) +
                           ((k > 0) ? A[c - nx * ny].front * x[c - nx * ny] : 0.0) +
                           ((k + 1 < nz) ? A[c].front * x[c + nx * ny] : 0.0);
                x[c] = (1.0 - w) * x[c] + w * (b[c] - r) / A[c].center;
            }
        }
}

int jeto_fdm_relax_red_black(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b,
                             double sor, uint32_t iterations, double* x) {
    for (uint32_t it = 0; it < iterations; ++it) {
        relax_colour((ssz)nx, (ssz)ny, (ssz)nz, A, b, sor, x, 0);
        relax_colour((ssz)nx, (ssz)ny, (ssz)nz, A, b, sor, x, 1);
    }
    return 0;
}

/* FdmGaussSeidelSolver3::relax — src/jet/fdm_gauss_seidel_solver3.cpp:103-120: the natural (lexicographic) ordering,
 * A.forEachIndex = i fastest, then j, then k, in place; the default ordering of FdmMgSolver3 / FdmMgpcgSolver3
 * (useRedBlackOrdering = false, include/jet/fdm_mgpcg_solver3.h:41).  Serial by construction. */
int jeto_fdm_relax(uint64_t unx, uint64_t uny, uint64_t unz, const jeto_row* A, const double* b, double w,
                   uint32_t iterations, double* x) {
    const ssz nx = (ssz)unx, ny = (ssz)uny, nz = (ssz)unz;
    for (uint32_t it = 0; it < iterations; ++it)
        for (ssz k = 0; k < nz; ++k)
            for (ssz j = 0; j < ny; ++j)
                for (ssz i = 0; i < nx; ++i) {
                    const ssz c = (k * ny + j) * nx + i;
                    double r = ((i > 0) ? A[c - 1].right * x[c - 1] : 0.0) +
                               ((i + 1 < nx) ? A[c].right * x[c + 1] : 0.0) +
                               ((j > 0) ? A[c - nx].up * x[c - nx] : 0.0) +
                               ((j + 1 < ny) ? A[c].up * x[c + nx] : 0.0) +
                               ((k > 0) ? A[c - nx * ny].front * x[c - nx * ny] : 0.0) +
                               ((k + 1 < nz) ? A[c].front * x[c + nx * ny] : 0.0);
                    x[c] = (1.0 - w) * x[c] + w * (b[c] - r) / A[c].center;
                }
    return 0;
}

/* the smoother FdmMgSolver3's constructor installs (src/jet/fdm_mg_solver3.cpp:26-49) */
static void mg_relax(const jeto_mgpcg_params* p, uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b,
                     uint32_t iterations, double* x) {
    if (p->use_red_black) jeto_fdm_relax_red_black(nx, ny, nz, A, b, p->sor_factor, iterations, x);
    else jeto_fdm_relax(nx, ny, nz, A, b, p->sor_factor, iterations, x);
}

/* FdmMgUtils3::restrict — src/jet/fdm_mg_linear_system3.cpp:43-98: (1,3,3,1)/8 tensor filter, edge-clamped,
 * accumulated z-outer / y / x-inner. (nx,ny,nz) is the FINER size. */
int jeto_fdm_restrict(uint64_t fnx, uint64_t fny, uint64_t fnz, const double* finer, double* coarser) {
    static const double kern[4] = {0.125, 0.375, 0.375, 0.125};
    ssz nx = (ssz)fnx / 2, ny = (ssz)fny / 2, nz = (ssz)fnz / 2;
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k) {
        ssz kI[4] = {(k > 0) ? 2 * k - 1 : 2 * k, 2 * k, 2 * k + 1, (k + 1 < nz) ? 2 * k + 2 : 2 * k + 1};
        for (ssz j = 0; j < ny; ++j) {
            ssz jI[4] = {(j > 0) ? 2 * j - 1 : 2 * j, 2 * j, 2 * j + 1, (j + 1 < ny) ? 2 * j + 2 : 2 * j + 1};
            for (ssz i = 0; i < nx; ++i) {
                ssz iI[4] = {(i > 0) ? 2 * i - 1 : 2 * i, 2 * i, 2 * i + 1, (i + 1 < nx) ? 2 * i + 2 : 2 * i + 1};
                double sum = 0.0;
                for (int z = 0; z < 4; ++z)
                    for (int y = 0; y < 4; ++y)
                        for (int x = 0; x < 4; ++x) {
                            double w = kern[x] * kern[y] * kern[z];
                            sum += w * finer[IDX(iI[x], jI[y], kI[z], fnx, fny)];
                        }
                coarser[IDX(i, j, k, nx, ny)] = sum;
            }
        }
    }
    return 0;
}

/* FdmMgUtils3::correct — src/jet/fdm_mg_linear_system3.cpp:100-179: finer += trilinear (1/4,3/4) prolongation.
 * The reference clamps k against n.y (:159): identical to clamping against n.z whenever ny == nz — the only shapes on which
 * the reference is well defined (for ny > nz it reads one coarse plane beyond the array, for ny < nz it clamps too early) and
 * the only shapes on which this function is compared with it.  For ny != nz the clamp is against n.z, the evident intent:
 * there is no reference answer to match there, only the requirement that the V-cycle stays a valid preconditioner. */
int jeto_fdm_correct(uint64_t fnx, uint64_t fny, uint64_t fnz, const double* coarser, double* finer) {
    ssz nx = (ssz)fnx, ny = (ssz)fny, nz = (ssz)fnz;
    ssz cnx = nx / 2, cny = ny / 2;
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx; ++i) {
                ssz iI[2], jI[2], kI[2];
                double iW[2], jW[2], kW[2];
                ssz ci = i / 2, cj = j / 2, ck = k / 2;
                if (i % 2 == 0) { iI[0] = (i > 1) ? ci - 1 : ci; iI[1] = ci; iW[0] = 0.25; iW[1] = 0.75; }
                else { iI[0] = ci; iI[1] = (i + 1 < nx) ? ci + 1 : ci; iW[0] = 0.75; iW[1] = 0.25; }
                if (j % 2 == 0) { jI[0] = (j > 1) ? cj - 1 : cj; jI[1] = cj; jW[0] = 0.25; jW[1] = 0.75; }
                else { jI[0] = cj; jI[1] = (j + 1 < ny) ? cj + 1 : cj; jW[0] = 0.75; jW[1] = 0.25; }
                if (k % 2 == 0) { kI[0] = (k > 1) ? ck - 1 : ck; kI[1] = ck; kW[0] = 0.25; kW[1] = 0.75; }
                else { kI[0] = ck; kI[1] = (k + 1 < nz) ? ck + 1 : ck; kW[0] = 0.75; kW[1] = 0.25; }
                size_t f = IDX(i, j, k, nx, ny);
                for (int z = 0; z < 2; ++z)
                    for (int y = 0; y < 2; ++y)
                        for (int x = 0; x < 2; ++x) {
                            double w = iW[x] * jW[y] * kW[z] * coarser[IDX(iI[x], jI[y], kI[z], cnx, cny)];
                            finer[f] += w;
                        }
            }
    return 0;
}

/* FdmMgUtils3::resizeArrayWithFinest — include/jet/detail/fdm_mg_linear_system3-inl.h:34-50. */
uint32_t jeto_mg_num_levels(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t max_levels) {
    uint32_t i = 1;
    for (; i < max_levels; ++i) {
        if (nx % 2 == 0 && ny % 2 == 0 && nz % 2 == 0) { nx >>= 1; ny >>= 1; nz >>= 1; }
        else break;
    }
    return i;
}

/* ====================================================================== MG V-cycle and PCG */

typedef struct {
    uint32_t levels;
    ssz nx[16], ny[16], nz[16];
    const jeto_row* A[16];
    double* x[16];
    double* b[16];
    double* buf[16];
} mg_ws;

static size_t lvl_n(const mg_ws* w, uint32_t l) { return (size_t)w->nx[l] * w->ny[l] * w->nz[l]; }

static void mg_alloc(mg_ws* w, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A) {
    w->levels = levels;
    for (uint32_t l = 0; l < levels; ++l) {
        w->nx[l] = (ssz)(nx >> l); w->ny[l] = (ssz)(ny >> l); w->nz[l] = (ssz)(nz >> l);
        w->A[l] = A[l];
        size_t n = lvl_n(w, l);
        w->x[l] = (double*)calloc(n, sizeof(double));
        w->b[l] = (double*)calloc(n, sizeof(double));
        w->buf[l] = (double*)calloc(n, sizeof(double));
    }
}

static void mg_free(mg_ws* w) {
    for (uint32_t l = 0; l < w->levels; ++l) { free(w->x[l]); free(w->b[l]); free(w->buf[l]); }
}

/* internal::mgVCycle — include/jet/detail/mg-inl.h:16-68 with the relax lambdas of FdmMgSolver3
 * (src/jet/fdm_mg_solver3.cpp:26-49: maxTolerance/buffer unused).  The trailing residual + l2Norm on every
 * level only produce MgResult.lastResidualNorm, which FdmMgpcgSolver3::Preconditioner discards; they do not
 * change x and are omitted (the finest-level norm is available through jeto_fdm_residual). */
static void vcycle(mg_ws* w, const jeto_mgpcg_params* p, uint32_t l) {
    mg_relax(p, w->nx[l], w->ny[l], w->nz[l], w->A[l], w->b[l], p->n_restriction, w->x[l]);
    if (l + 1 < w->levels) {
        jeto_fdm_residual(w->nx[l], w->ny[l], w->nz[l], w->A[l], w->x[l], w->b[l], w->buf[l]);
        jeto_fdm_restrict(w->nx[l], w->ny[l], w->nz[l], w->buf[l], w->b[l + 1]);
        memset(w->x[l + 1], 0, lvl_n(w, l + 1) * sizeof(double));
        vcycle(w, p, l + 1);
        jeto_fdm_correct(w->nx[l], w->ny[l], w->nz[l], w->x[l + 1], w->x[l]);
        mg_relax(p, w->nx[l], w->ny[l], w->nz[l], w->A[l], w->b[l], l > 0 ? p->n_correction : p->n_final, w->x[l]);
    } else {
        mg_relax(p, w->nx[l], w->ny[l], w->nz[l], w->A[l], w->b[l], p->n_coarsest, w->x[l]);
    }
}

int jeto_fdm_vcycle(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A,
                    const double* b, double* x, const jeto_mgpcg_params* p) {
    if (levels > 16) return 3;
    mg_ws w;
    mg_alloc(&w, nx, ny, nz, levels, A);
    size_t n = lvl_n(&w, 0);
    memcpy(w.x[0], x, n * sizeof(double));
    memcpy(w.b[0], b, n * sizeof(double));
    vcycle(&w, p, 0);
    memcpy(x, w.x[0], n * sizeof(double));
    mg_free(&w);
    return 0;
}

/* FdmMgpcgSolver3::Preconditioner::solve — src/jet/fdm_mgpcg_solver3.cpp:21-36: the V-cycle starts from the
 * CURRENT content of the output vector (not zero) with b = input on the finest level. */
static void precond(mg_ws* w, const jeto_mgpcg_params* p, const double* b, double* x) {
    size_t n = lvl_n(w, 0);
    memcpy(w->x[0], x, n * sizeof(double));
    memcpy(w->b[0], b, n * sizeof(double));
    vcycle(w, p, 0);
    memcpy(x, w->x[0], n * sizeof(double));
}

/* pcg<FdmBlas3, Preconditioner> — include/jet/detail/cg-inl.h:18-94, called as in
 * FdmMgpcgSolver3::solve — src/jet/fdm_mgpcg_solver3.cpp:54-80 (x zeroed first). */
static int mgpcg_core(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A,
                      const double* b, double* x, const jeto_mgpcg_params* p, jeto_mgpcg_result* result) {
    if (levels > 16) return 3;
    mg_ws w;
    mg_alloc(&w, nx, ny, nz, levels, A);
    size_t n = lvl_n(&w, 0);
    double* r = (double*)calloc(n, sizeof(double));
    double* d = (double*)calloc(n, sizeof(double));
    double* q = (double*)calloc(n, sizeof(double));
    double* s = (double*)calloc(n, sizeof(double));
    double t0 = now_ms();
    memset(x, 0, n * sizeof(double));

    jeto_fdm_residual(nx, ny, nz, A[0], x, b, r);
    precond(&w, p, r, d);
    double sigmaNew;
    jeto_fdm_dot(n, r, d, &sigmaNew);

    uint32_t iter = 0;
    int trigger = 0;
    while (sigmaNew > p->tolerance * p->tolerance && iter < p->max_iterations) {
        jeto_fdm_mvm(nx, ny, nz, A[0], d, q);
        double dq;
        jeto_fdm_dot(n, d, q, &dq);
        double alpha = sigmaNew / dq;
        jeto_fdm_axpy(n, alpha, d, x, x);
        if (trigger || (iter % 50 == 0 && iter > 0)) {
            jeto_fdm_residual(nx, ny, nz, A[0], x, b, r);
            trigger = 0;
        } else {
            jeto_fdm_axpy(n, -alpha, q, r, r);
        }
        precond(&w, p, r, s);
        double sigmaOld = sigmaNew;
        jeto_fdm_dot(n, r, s, &sigmaNew);
        if (sigmaNew > sigmaOld) trigger = 1;
        double beta = sigmaNew / sigmaOld;
        jeto_fdm_axpy(n, beta, d, s, d);
        ++iter;
    }
    double t1 = now_ms();
    if (result) {
        result->iterations = iter;
        result->last_residual = sqrt(fabs(sigmaNew));
        result->returned = (result->last_residual <= p->tolerance || iter < p->max_iterations) ? 1 : 0;
        jeto_fdm_residual(nx, ny, nz, A[0], x, b, r);
        double rr, bb;
        jeto_fdm_dot(n, r, r, &rr);
        jeto_fdm_dot(n, b, b, &bb);
        result->true_residual_l2 = sqrt(rr);
        result->b_l2 = sqrt(bb);
        result->solve_ms = t1 - t0;
    }
    free(r); free(d); free(q); free(s);
    mg_free(&w);
    return 0;
}

int jeto_mgpcg_solve(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A,
                     const double* b, double* x, const jeto_mgpcg_params* p, jeto_mgpcg_result* result) {
    return mgpcg_core(nx, ny, nz, levels, A, b, x, p, result);
}

/* ====================================================================== ICCG */
/* FdmIccgSolver3::Preconditioner::build — src/jet/fdm_iccg_solver3.cpp:14-36 (lexicographic recurrence). */
static void ic_build(ssz nx, ssz ny, ssz nz, const jeto_row* A, double* d) {
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx; ++i) {
                size_t c = IDX(i, j, k, nx, ny);
                double denom = A[c].center - ((i > 0) ? (A[c - 1].right * A[c - 1].right) * d[c - 1] : 0.0) -
                               ((j > 0) ? (A[c - nx].up * A[c - nx].up) * d[c - nx] : 0.0) -
                               ((k > 0) ? (A[c - nx * ny].front * A[c - nx * ny].front) * d[c - nx * ny] : 0.0);
                d[c] = (fabs(denom) > 0.0) ? 1.0 / denom : 0.0;
            }
}

/* FdmIccgSolver3::Preconditioner::solve — :38-70: forward substitution into y, backward into x. */
static void ic_solve(ssz nx, ssz ny, ssz nz, const jeto_row* A, const double* d, const double* b, double* y, double* x) {
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx; ++i) {
                size_t c = IDX(i, j, k, nx, ny);
                y[c] = (b[c] - ((i > 0) ? A[c - 1].right * y[c - 1] : 0.0) - ((j > 0) ? A[c - nx].up * y[c - nx] : 0.0) -
                        ((k > 0) ? A[c - nx * ny].front * y[c - nx * ny] : 0.0)) * d[c];
            }
    for (ssz k = nz - 1; k >= 0; --k)
        for (ssz j = ny - 1; j >= 0; --j)
            for (ssz i = nx - 1; i >= 0; --i) {
                size_t c = IDX(i, j, k, nx, ny);
                x[c] = (y[c] - ((i + 1 < nx) ? A[c].right * x[c + 1] : 0.0) - ((j + 1 < ny) ? A[c].up * x[c + nx] : 0.0) -
                        ((k + 1 < nz) ? A[c].front * x[c + nx * ny] : 0.0)) * d[c];
            }
}

int jeto_iccg_precondition(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b, double* d, double* y,
                           double* x) {
    size_t n = (size_t)nx * ny * nz;
    double* dd = d ? d : (double*)calloc(n, sizeof(double));
    ic_build((ssz)nx, (ssz)ny, (ssz)nz, A, dd);
    ic_solve((ssz)nx, (ssz)ny, (ssz)nz, A, dd, b, y, x);
    if (!d) free(dd);
    return 0;
}

/* FdmIccgSolver3::solve(FdmLinearSystem3*) — :152-188 = pcg<FdmBlas3, Preconditioner> (cg-inl.h:18-94) with x = 0. */
int jeto_iccg_solve(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b, double* x,
                    uint32_t max_iterations, double tolerance, jeto_mgpcg_result* result) {
    size_t n = (size_t)nx * ny * nz;
    double* r = (double*)calloc(n, sizeof(double));
    double* d = (double*)calloc(n, sizeof(double));
    double* q = (double*)calloc(n, sizeof(double));
    double* s = (double*)calloc(n, sizeof(double));
    double* dinv = (double*)calloc(n, sizeof(double));
    double* y = (double*)calloc(n, sizeof(double));
    double t0 = now_ms();
    memset(x, 0, n * sizeof(double));
    ic_build((ssz)nx, (ssz)ny, (ssz)nz, A, dinv);
    jeto_fdm_residual(nx, ny, nz, A, x, b, r);
    ic_solve((ssz)nx, (ssz)ny, (ssz)nz, A, dinv, r, y, d);
    double sigmaNew;
    jeto_fdm_dot(n, r, d, &sigmaNew);
    uint32_t iter = 0;
    int trigger = 0;
    while (sigmaNew > tolerance * tolerance && iter < max_iterations) {
        jeto_fdm_mvm(nx, ny, nz, A, d, q);
        double dq;
        jeto_fdm_dot(n, d, q, &dq);
        double alpha = sigmaNew / dq;
        jeto_fdm_axpy(n, alpha, d, x, x);
        if (trigger || (iter % 50 == 0 && iter > 0)) {
            jeto_fdm_residual(nx, ny, nz, A, x, b, r);
            trigger = 0;
        } else {
            jeto_fdm_axpy(n, -alpha, q, r, r);
        }
        ic_solve((ssz)nx, (ssz)ny, (ssz)nz, A, dinv, r, y, s);
        double sigmaOld = sigmaNew;
        jeto_fdm_dot(n, r, s, &sigmaNew);
        if (sigmaNew > sigmaOld) trigger = 1;
        double beta = sigmaNew / sigmaOld;
        jeto_fdm_axpy(n, beta, d, s, d);
        ++iter;
    }
    double t1 = now_ms();
    if (result) {
        result->iterations = iter;
        result->last_residual = sqrt(fabs(sigmaNew));
        result->returned = (result->last_residual <= tolerance || iter < max_iterations) ? 1 : 0;
        jeto_fdm_residual(nx, ny, nz, A, x, b, r);
        double rr, bb;
        jeto_fdm_dot(n, r, r, &rr);
        jeto_fdm_dot(n, b, b, &bb);
        result->true_residual_l2 = sqrt(rr);
        result->b_l2 = sqrt(bb);
        result->solve_ms = t1 - t0;
    }
    free(r); free(d); free(q); free(s); free(dinv); free(y);
    return 0;
}

/* ====================================================================== samplers */

/* getBarycentric — include/jet/detail/math_utils-inl.h:146-174 (iLow is always 0 on this path). */
static inline void barycentric(double x, ssz iHigh, ssz* i, double* f) {
    double s = floor(x);
    *i = (ssz)s;
    if (iHigh == 0) { *i = 0; *f = 0; }
    else if (*i < 0) { *i = 0; *f = 0; }
    else if (*i > iHigh - 1) { *i = iHigh - 1; *f = 1; }
    else { *f = x - s; }
}

static inline void barycentricf(float x, ssz iHigh, ssz* i, float* f) {
    float s = floorf(x);
    *i = (ssz)s;
    if (iHigh == 0) { *i = 0; *f = 0; }
    else if (*i < 0) { *i = 0; *f = 0; }
    else if (*i > iHigh - 1) { *i = iHigh - 1; *f = 1; }
    else { *f = x - s; }
}

static inline double lerpd(double a, double b, double f) { return (1 - f) * a + f * b; }
static inline float lerpf(float a, float b, float f) { return (1 - f) * a + f * b; }
static inline ssz minz(ssz a, ssz b) { return a < b ? a : b; }

typedef struct {
    const double* data;
    ssz nx, ny, nz;
    double hx, hy, hz, ox, oy, oz;
} sampler3;

/* LinearArraySampler3<double,double>::operator() — include/jet/detail/array_samplers3-inl.h:115-149:
 * normalises with a DIVISION by h, then nested lerps x -> y -> z (math_utils-inl.h:176-211). */
static double sample3(const sampler3* s, double px, double py, double pz) {
    ssz i, j, k;
    double fx, fy, fz;
    barycentric((px - s->ox) / s->hx, s->nx - 1, &i, &fx);
    barycentric((py - s->oy) / s->hy, s->ny - 1, &j, &fy);
    barycentric((pz - s->oz) / s->hz, s->nz - 1, &k, &fz);
    ssz ip1 = minz(i + 1, s->nx - 1), jp1 = minz(j + 1, s->ny - 1), kp1 = minz(k + 1, s->nz - 1);
    const double* d = s->data;
#define S3(a, b, c) d[IDX(a, b, c, s->nx, s->ny)]
    double r = lerpd(lerpd(lerpd(S3(i, j, k), S3(ip1, j, k), fx), lerpd(S3(i, jp1, k), S3(ip1, jp1, k), fx), fy),
                     lerpd(lerpd(S3(i, j, kp1), S3(ip1, j, kp1), fx), lerpd(S3(i, jp1, kp1), S3(ip1, jp1, kp1), fx), fy),
                     fz);
#undef S3
    return r;
}

/* LinearArraySampler3::getCoordinatesAndWeights — array_samplers3-inl.h:151-193: normalises with a
 * MULTIPLICATION by invH = 1/h; weights are explicit triple products. */
static void coords_weights(const sampler3* s, double px, double py, double pz, size_t idx[8], double w[8]) {
    ssz i, j, k;
    double fx, fy, fz;
    double ihx = 1.0 / s->hx, ihy = 1.0 / s->hy, ihz = 1.0 / s->hz;
    barycentric((px - s->ox) * ihx, s->nx - 1, &i, &fx);
    barycentric((py - s->oy) * ihy, s->ny - 1, &j, &fy);
    barycentric((pz - s->oz) * ihz, s->nz - 1, &k, &fz);
    ssz ip1 = minz(i + 1, s->nx - 1), jp1 = minz(j + 1, s->ny - 1), kp1 = minz(k + 1, s->nz - 1);
    idx[0] = IDX(i, j, k, s->nx, s->ny);     idx[1] = IDX(ip1, j, k, s->nx, s->ny);
    idx[2] = IDX(i, jp1, k, s->nx, s->ny);   idx[3] = IDX(ip1, jp1, k, s->nx, s->ny);
    idx[4] = IDX(i, j, kp1, s->nx, s->ny);   idx[5] = IDX(ip1, j, kp1, s->nx, s->ny);
    idx[6] = IDX(i, jp1, kp1, s->nx, s->ny); idx[7] = IDX(ip1, jp1, kp1, s->nx, s->ny);
    w[0] = (1 - fx) * (1 - fy) * (1 - fz); w[1] = fx * (1 - fy) * (1 - fz);
    w[2] = (1 - fx) * fy * (1 - fz);       w[3] = fx * fy * (1 - fz);
    w[4] = (1 - fx) * (1 - fy) * fz;       w[5] = fx * (1 - fy) * fz;
    w[6] = (1 - fx) * fy * fz;             w[7] = fx * fy * fz;
}

/* ====================================================================== level-set helpers */

/* fractionInsideSdf — include/jet/detail/level_set_utils-inl.h:52-63. */
static double frac_inside_sdf(double p0, double p1) {
    if (p0 < 0 && p1 < 0) return 1;
    if (p0 < 0 && !(p1 < 0)) return p0 / (p0 - p1);
    if (!(p0 < 0) && p1 < 0) return p1 / (p1 - p0);
    return 0;
}

static void cycle4(double* a) { double t = a[0]; a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = t; }

/* fractionInside (marching-squares area) — level_set_utils-inl.h:72-136.  `0.5f * x` with x double is exact. */
static double frac_inside4(double bl, double br, double tl, double tr) {
    int cnt = (bl < 0) + (tl < 0) + (br < 0) + (tr < 0);
    double l[4] = {bl, br, tr, tl};
    if (cnt == 4) return 1;
    if (cnt == 3) {
        while (l[0] < 0) cycle4(l);
        double s0 = 1 - frac_inside_sdf(l[0], l[3]);
        double s1 = 1 - frac_inside_sdf(l[0], l[1]);
        return 1 - 0.5 * s0 * s1;
    }
    if (cnt == 2) {
        while (l[0] >= 0 || !(l[1] < 0 || l[2] < 0)) cycle4(l);
        if (l[1] < 0) {
            double sl = frac_inside_sdf(l[0], l[3]);
            double sr = frac_inside_sdf(l[1], l[2]);
            return 0.5 * (sl + sr);
        }
        double mid = 0.25 * (l[0] + l[1] + l[2] + l[3]);
        if (mid < 0) {
            double area = 0;
            double s1 = 1 - frac_inside_sdf(l[0], l[3]);
            double s3 = 1 - frac_inside_sdf(l[2], l[3]);
            area += 0.5 * s1 * s3;
            double s2 = 1 - frac_inside_sdf(l[2], l[1]);
            double s0 = 1 - frac_inside_sdf(l[0], l[1]);
            area += 0.5 * s0 * s2;
            return 1 - area;
        } else {
            double area = 0;
            double s0 = frac_inside_sdf(l[0], l[1]);
            double s1 = frac_inside_sdf(l[0], l[3]);
            area += 0.5 * s0 * s1;
            double s2 = frac_inside_sdf(l[2], l[1]);
            double s3 = frac_inside_sdf(l[2], l[3]);
            area += 0.5 * s2 * s3;
            return area;
        }
    }
    if (cnt == 1) {
        while (l[0] >= 0) cycle4(l);
        double s0 = frac_inside_sdf(l[0], l[3]);
        double s1 = frac_inside_sdf(l[0], l[1]);
        return 0.5 * s0 * s1;
    }
    return 0;
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ====================================================================== fractional pressure solver */

typedef struct {
    uint32_t levels;
    ssz nx[16], ny[16], nz[16];
    float* sdf[16];
    float* uw[16];
    float* vw[16];
    float* ww[16];
    jeto_row* A[16];
    double* b0;
    double* x0;
} psys;

static psys g_psys; /* last system, for jeto_pressure_get_system */

static void psys_free(psys* s) {
    for (uint32_t l = 0; l < s->levels; ++l) { free(s->sdf[l]); free(s->uw[l]); free(s->vw[l]); free(s->ww[l]); free(s->A[l]); }
    free(s->b0); free(s->x0);
    memset(s, 0, sizeof(*s));
}

/* Anonymous-namespace restrict(Array3<float>) — grid_fractional_single_phase_pressure_solver3.cpp:29-115.
 * An axis whose finer size != 2*coarser size is the face grid's staggered axis: 3-tap {0,1,0} kernel. */
static void restrict_f32(const float* finer, ssz fx, ssz fy, ssz fz, float* coarser, ssz cx, ssz cy, ssz cz) {
    static const float centered[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    static const float staggered[4] = {0.f, 1.f, 0.f, 0.f};
    int ks[3] = {fx != 2 * cx ? 3 : 4, fy != 2 * cy ? 3 : 4, fz != 2 * cz ? 3 : 4};
    const float* kn[3] = {ks[0] == 3 ? staggered : centered, ks[1] == 3 ? staggered : centered,
                          ks[2] == 3 ? staggered : centered};
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < cz; ++k) {
        ssz kI[4], jI[4], iI[4];
        if (ks[2] == 3) { kI[0] = (k > 0) ? 2 * k - 1 : 2 * k; kI[1] = 2 * k; kI[2] = (k + 1 < cz) ? 2 * k + 1 : 2 * k; kI[3] = 0; }
        else { kI[0] = (k > 0) ? 2 * k - 1 : 2 * k; kI[1] = 2 * k; kI[2] = 2 * k + 1; kI[3] = (k + 1 < cz) ? 2 * k + 2 : 2 * k + 1; }
        for (ssz j = 0; j < cy; ++j) {
            if (ks[1] == 3) { jI[0] = (j > 0) ? 2 * j - 1 : 2 * j; jI[1] = 2 * j; jI[2] = (j + 1 < cy) ? 2 * j + 1 : 2 * j; jI[3] = 0; }
            else { jI[0] = (j > 0) ? 2 * j - 1 : 2 * j; jI[1] = 2 * j; jI[2] = 2 * j + 1; jI[3] = (j + 1 < cy) ? 2 * j + 2 : 2 * j + 1; }
            for (ssz i = 0; i < cx; ++i) {
                if (ks[0] == 3) { iI[0] = (i > 0) ? 2 * i - 1 : 2 * i; iI[1] = 2 * i; iI[2] = (i + 1 < cx) ? 2 * i + 1 : 2 * i; iI[3] = 0; }
                else { iI[0] = (i > 0) ? 2 * i - 1 : 2 * i; iI[1] = 2 * i; iI[2] = 2 * i + 1; iI[3] = (i + 1 < cx) ? 2 * i + 2 : 2 * i + 1; }
                float sum = 0.0f;
                for (int z = 0; z < ks[2]; ++z)
                    for (int y = 0; y < ks[1]; ++y)
                        for (int x = 0; x < ks[0]; ++x) {
                            float w = kn[0][x] * kn[1][y] * kn[2][z];
                            sum += w * finer[IDX(iI[x], jI[y], kI[z], fx, fy)];
                        }
                coarser[IDX(i, j, k, cx, cy)] = sum;
            }
        }
    }
}

static double weight_from_frac(double frac) {
    /* buildWeights — ...pressure_solver3.cpp:563-571 (kMinWeight = 0.01). */
    double weight = clampd(1.0 - frac, 0.0, 1.0);
    if (weight < 0.01 && weight > 0.0) weight = 0.01;
    return weight;
}

/* buildSingleSystem (uncompressed) — ...pressure_solver3.cpp:117-267.  `u,v,w` may be NULL (coarse levels):
 * b is then left zero — it is dead in both MG solvers (the V-cycle overwrites b[l>0]). */
static void build_single_system(jeto_row* A, double* b, ssz nx, ssz ny, ssz nz, const float* sdf, const float* uw,
                                const float* vw, const float* ww, const double* u, const double* v, const double* w,
                                const double* bu, const double* bv, const double* bw, double hx, double hy, double hz) {
    double ihx = 1.0 / hx, ihy = 1.0 / hy, ihz = 1.0 / hz;
    double ihx2 = ihx * ihx, ihy2 = ihy * ihy, ihz2 = ihz * ihz;
    ssz unx = nx + 1, vny = ny + 1;
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx; ++i) {
                size_t c = IDX(i, j, k, nx, ny);
                jeto_row row = {0.0, 0.0, 0.0, 0.0};
                double bb = 0.0;
                double cphi = sdf[c];
                size_t u1 = IDX(i + 1, j, k, unx, ny), u0 = IDX(i, j, k, unx, ny);
                size_t v1 = IDX(i, j + 1, k, nx, vny), v0 = IDX(i, j, k, nx, vny);
                size_t w1 = IDX(i, j, k + 1, nx, ny), w0 = IDX(i, j, k, nx, ny);
                if (cphi < 0) {
                    double term, theta;
#define U(idx) (u ? u[idx] : 0.0)
#define V(idx) (v ? v[idx] : 0.0)
#define W(idx) (w ? w[idx] : 0.0)
                    if (i + 1 < nx) {
                        term = uw[u1] * ihx2;
                        double nphi = sdf[c + 1];
                        if (nphi < 0) { row.center += term; row.right -= term; }
                        else { theta = fmax(frac_inside_sdf(cphi, nphi), 0.01); row.center += term / theta; }
                        bb += uw[u1] * U(u1) * ihx;
                    } else { bb += U(u1) * ihx; }
                    if (i > 0) {
                        term = uw[u0] * ihx2;
                        double nphi = sdf[c - 1];
                        if (nphi < 0) { row.center += term; }
                        else { theta = fmax(frac_inside_sdf(cphi, nphi), 0.01); row.center += term / theta; }
                        bb -= uw[u0] * U(u0) * ihx;
                    } else { bb -= U(u0) * ihx; }
                    if (j + 1 < ny) {
                        term = vw[v1] * ihy2;
                        double nphi = sdf[c + nx];
                        if (nphi < 0) { row.center += term; row.up -= term; }
                        else { theta = fmax(frac_inside_sdf(cphi, nphi), 0.01); row.center += term / theta; }
                        bb += vw[v1] * V(v1) * ihy;
                    } else { bb += V(v1) * ihy; }
                    if (j > 0) {
                        term = vw[v0] * ihy2;
                        double nphi = sdf[c - nx];
                        if (nphi < 0) { row.center += term; }
                        else { theta = fmax(frac_inside_sdf(cphi, nphi), 0.01); row.center += term / theta; }
                        bb -= vw[v0] * V(v0) * ihy;
                    } else { bb -= V(v0) * ihy; }
                    if (k + 1 < nz) {
                        term = ww[w1] * ihz2;
                        double nphi = sdf[c + nx * ny];
                        if (nphi < 0) { row.center += term; row.front -= term; }
                        else { theta = fmax(frac_inside_sdf(cphi, nphi), 0.01); row.center += term / theta; }
                        bb += ww[w1] * W(w1) * ihz;
                    } else { bb += W(w1) * ihz; }
                    if (k > 0) {
                        term = ww[w0] * ihz2;
                        double nphi = sdf[c - nx * ny];
                        if (nphi < 0) { row.center += term; }
                        else { theta = fmax(frac_inside_sdf(cphi, nphi), 0.01); row.center += term / theta; }
                        bb -= ww[w0] * W(w0) * ihz;
                    } else { bb -= W(w0) * ihz; }
#undef U
#undef V
#undef W
                    /* moving-boundary term, :242-255 */
                    double bc = (1.0 - uw[u1]) * (bu ? bu[u1] : 0.0) * ihx - (1.0 - uw[u0]) * (bu ? bu[u0] : 0.0) * ihx +
                                (1.0 - vw[v1]) * (bv ? bv[v1] : 0.0) * ihy - (1.0 - vw[v0]) * (bv ? bv[v0] : 0.0) * ihy +
                                (1.0 - ww[w1]) * (bw ? bw[w1] : 0.0) * ihz - (1.0 - ww[w0]) * (bw ? bw[w0] : 0.0) * ihz;
                    bb += bc;
                    if (row.center < DBL_EPSILON) { row.center = 1.0; bb = 0.0; }
                } else {
                    row.center = 1.0;
                }
                A[c] = row;
                if (b) b[c] = bb;
            }
}

/* GridFractionalSinglePhasePressureSolver3::solve with an MGPCG linear solver —
 * ...pressure_solver3.cpp:456-484 (buildWeights :520-636, buildSystem :651-713, applyPressureGradient :715-765). */
static int pressure_core(const jeto_grid* g, const double* u_in, const double* v_in, const double* w_in,
                         const double* collider_sdf, const double* bu, const double* bv, const double* bw,
                         const double* fluid_sdf, const jeto_mgpcg_params* p, double* u_out, double* v_out,
                         double* w_out, double* pressure_out, jeto_mgpcg_result* result) {
    psys* S = &g_psys;
    psys_free(S);
    ssz nx = (ssz)g->nx, ny = (ssz)g->ny, nz = (ssz)g->nz;
    double hx = g->hx, hy = g->hy, hz = g->hz;
    S->levels = jeto_mg_num_levels(nx, ny, nz, p->max_levels);
    for (uint32_t l = 0; l < S->levels; ++l) {
        ssz lx = nx >> l, ly = ny >> l, lz = nz >> l;
        S->nx[l] = lx; S->ny[l] = ly; S->nz[l] = lz;
        S->sdf[l] = (float*)calloc((size_t)lx * ly * lz, sizeof(float));
        S->uw[l] = (float*)calloc((size_t)(lx + 1) * ly * lz, sizeof(float));
        S->vw[l] = (float*)calloc((size_t)lx * (ly + 1) * lz, sizeof(float));
        S->ww[l] = (float*)calloc((size_t)lx * ly * (lz + 1), sizeof(float));
        S->A[l] = (jeto_row*)calloc((size_t)lx * ly * lz, sizeof(jeto_row));
    }
    size_t n = (size_t)nx * ny * nz;
    S->b0 = (double*)calloc(n, sizeof(double));
    S->x0 = (double*)calloc(n, sizeof(double));

    /* cell-centred grid samplers: dataOrigin = origin + 0.5*h (src/jet/cell_centered_scalar_grid3.cpp:60-62) */
    sampler3 fs = {fluid_sdf, nx, ny, nz, hx, hy, hz, g->ox + 0.5 * hx, g->oy + 0.5 * hy, g->oz + 0.5 * hz};
    sampler3 cs = fs; cs.data = collider_sdf;

    /* _fluidSdf[0] = float(fluidSdf.sample(cellCenterPosition)); cellCenter = o + h*(i+0.5) (src/jet/grid3.cpp:32-38) */
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx; ++i) {
                double val = fluid_sdf ? sample3(&fs, g->ox + hx * (i + 0.5), g->oy + hy * (j + 0.5), g->oz + hz * (k + 0.5))
                                       : -DBL_MAX;
                S->sdf[0][IDX(i, j, k, nx, ny)] = (float)val;
            }
    /* face origins (src/jet/face_centered_grid3.cpp:442-444) and positions (:216-238) */
    double uox = g->ox + 0.5 * 0.0, uoy = g->oy + 0.5 * hy, uoz = g->oz + 0.5 * hz;
    double vox = g->ox + 0.5 * hx, voy = g->oy + 0.5 * 0.0, voz = g->oz + 0.5 * hz;
    double wox = g->ox + 0.5 * hx, woy = g->oy + 0.5 * hy, woz = g->oz + 0.5 * 0.0;
#define BS(px, py, pz) (collider_sdf ? sample3(&cs, (px), (py), (pz)) : DBL_MAX)
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx + 1; ++i) {
                double px = uox + hx * i, py = uoy + hy * j, pz = uoz + hz * k;
                double p0 = BS(px + 0.0, py + -0.5 * hy, pz + -0.5 * hz);
                double p1 = BS(px + 0.0, py + 0.5 * hy, pz + -0.5 * hz);
                double p2 = BS(px + 0.0, py + -0.5 * hy, pz + 0.5 * hz);
                double p3 = BS(px + 0.0, py + 0.5 * hy, pz + 0.5 * hz);
                S->uw[0][IDX(i, j, k, nx + 1, ny)] = (float)weight_from_frac(frac_inside4(p0, p1, p2, p3));
            }
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny + 1; ++j)
            for (ssz i = 0; i < nx; ++i) {
                double px = vox + hx * i, py = voy + hy * j, pz = voz + hz * k;
                double p0 = BS(px + -0.5 * hx, py + 0.0, pz + -0.5 * hz);
                double p1 = BS(px + -0.5 * hx, py + 0.0, pz + 0.5 * hz);
                double p2 = BS(px + 0.5 * hx, py + 0.0, pz + -0.5 * hz);
                double p3 = BS(px + 0.5 * hx, py + 0.0, pz + 0.5 * hz);
                S->vw[0][IDX(i, j, k, nx, ny + 1)] = (float)weight_from_frac(frac_inside4(p0, p1, p2, p3));
            }
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz + 1; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx; ++i) {
                double px = wox + hx * i, py = woy + hy * j, pz = woz + hz * k;
                double p0 = BS(px + -0.5 * hx, py + -0.5 * hy, pz + 0.0);
                double p1 = BS(px + -0.5 * hx, py + 0.5 * hy, pz + 0.0);
                double p2 = BS(px + 0.5 * hx, py + -0.5 * hy, pz + 0.0);
                double p3 = BS(px + 0.5 * hx, py + 0.5 * hy, pz + 0.0);
                S->ww[0][IDX(i, j, k, nx, ny)] = (float)weight_from_frac(frac_inside4(p0, p1, p2, p3));
            }
#undef BS
    for (uint32_t l = 1; l < S->levels; ++l) {
        ssz fx = S->nx[l - 1], fy = S->ny[l - 1], fz = S->nz[l - 1], cx = S->nx[l], cy = S->ny[l], cz = S->nz[l];
        restrict_f32(S->sdf[l - 1], fx, fy, fz, S->sdf[l], cx, cy, cz);
        restrict_f32(S->uw[l - 1], fx + 1, fy, fz, S->uw[l], cx + 1, cy, cz);
        restrict_f32(S->vw[l - 1], fx, fy + 1, fz, S->vw[l], cx, cy + 1, cz);
        restrict_f32(S->ww[l - 1], fx, fy, fz + 1, S->ww[l], cx, cy, cz + 1);
    }
    /* buildSystem: level l has spacing h * 2^l (:693-700) */
    double lhx = hx, lhy = hy, lhz = hz;
    for (uint32_t l = 0; l < S->levels; ++l) {
        build_single_system(S->A[l], l == 0 ? S->b0 : NULL, S->nx[l], S->ny[l], S->nz[l], S->sdf[l], S->uw[l],
                            S->vw[l], S->ww[l], l == 0 ? u_in : NULL, l == 0 ? v_in : NULL, l == 0 ? w_in : NULL,
                            bu, bv, bw, lhx, lhy, lhz);
        /* on coarse levels the reference samples boundaryVel at coarse face positions for a dead b: skipped */
        bu = bv = bw = NULL;
        lhx *= 2.0; lhy *= 2.0; lhz *= 2.0;
    }
    const jeto_row* Ac[16];
    for (uint32_t l = 0; l < S->levels; ++l) Ac[l] = S->A[l];
    /* FdmIccgSolver3 works on the level-0 system alone; with useCompressedLinearSystem the reference drops the air rows
     * (identity rows, b = 0, uncoupled): the port solves the uncompressed system in both cases (same solution on the liquid
     * cells to rounding; the unmodified reference covers the compressed arithmetic itself). */
    int rc = (p->linear_solver & 1) ? jeto_iccg_solve(nx, ny, nz, S->A[0], S->b0, S->x0, p->max_iterations, p->tolerance, result)
                                    : mgpcg_core(nx, ny, nz, S->levels, Ac, S->b0, S->x0, p, result);
    if (rc) return rc;

    /* applyPressureGradient — :715-765 */
    double ihx = 1.0 / hx, ihy = 1.0 / hy, ihz = 1.0 / hz;
    const float* sdf = S->sdf[0];
    const double* x = S->x0;
#pragma omp parallel for schedule(static)
    for (ssz k = 0; k < nz; ++k)
        for (ssz j = 0; j < ny; ++j)
            for (ssz i = 0; i < nx; ++i) {
                size_t c = IDX(i, j, k, nx, ny);
                double cphi = sdf[c];
                if (i + 1 < nx) {
                    size_t f = IDX(i + 1, j, k, nx + 1, ny);
                    if (S->uw[0][f] > 0.0 && (cphi < 0 || sdf[c + 1] < 0)) {
                        double theta = fmax(frac_inside_sdf(cphi, (double)sdf[c + 1]), 0.01);
                        u_out[f] = u_in[f] + ihx / theta * (x[c + 1] - x[c]);
                    }
                }
                if (j + 1 < ny) {
                    size_t f = IDX(i, j + 1, k, nx, ny + 1);
                    if (S->vw[0][f] > 0.0 && (cphi < 0 || sdf[c + nx] < 0)) {
                        double theta = fmax(frac_inside_sdf(cphi, (double)sdf[c + nx]), 0.01);
                        v_out[f] = v_in[f] + ihy / theta * (x[c + nx] - x[c]);
                    }
                }
                if (k + 1 < nz) {
                    size_t f = IDX(i, j, k + 1, nx, ny);
                    if (S->ww[0][f] > 0.0 && (cphi < 0 || sdf[c + nx * ny] < 0)) {
                        double theta = fmax(frac_inside_sdf(cphi, (double)sdf[c + nx * ny]), 0.01);
                        w_out[f] = w_in[f] + ihz / theta * (x[c + nx * ny] - x[c]);
                    }
                }
            }
    if (pressure_out) memcpy(pressure_out, x, n * sizeof(double));
    return 0;
}

int jeto_pressure_solve(const jeto_grid* g, const double* u_in, const double* v_in, const double* w_in,
                        const double* collider_sdf, const double* bu, const double* bv, const double* bw,
                        const double* fluid_sdf, const jeto_mgpcg_params* p, double* u_out, double* v_out,
                        double* w_out, double* pressure_out, jeto_mgpcg_result* result) {
    /* inputs may alias outputs (the reference's tests call solve(vel, ..., &vel)); every face is read and
     * written by the same loop iteration, so in-place is safe, but the system build must see the inputs first
     * — it does: the gradient is applied last. */
    return pressure_core(g, u_in, v_in, w_in, collider_sdf, bu, bv, bw, fluid_sdf, p, u_out, v_out, w_out,
                         pressure_out, result);
}

int jeto_pressure_get_system(uint32_t level, jeto_row* A, double* b, float* fluid_sdf, float* uw, float* vw,
                             float* ww) {
    psys* S = &g_psys;
    if (level >= S->levels) return 1;
    size_t lx = S->nx[level], ly = S->ny[level], lz = S->nz[level];
    if (A) memcpy(A, S->A[level], lx * ly * lz * sizeof(jeto_row));
    if (b) { if (level == 0) memcpy(b, S->b0, lx * ly * lz * sizeof(double)); else memset(b, 0, lx * ly * lz * sizeof(double)); }
    if (fluid_sdf) memcpy(fluid_sdf, S->sdf[level], lx * ly * lz * sizeof(float));
    if (uw) memcpy(uw, S->uw[level], (lx + 1) * ly * lz * sizeof(float));
    if (vw) memcpy(vw, S->vw[level], lx * (ly + 1) * lz * sizeof(float));
    if (ww) memcpy(ww, S->ww[level], lx * ly * (lz + 1) * sizeof(float));
    return 0;
}

#include "jet_port_collider.inc"
#include "jet_port_particles.inc"
#include "jet_port_diffusion.inc"
#include "jet_port_smoke.inc"
