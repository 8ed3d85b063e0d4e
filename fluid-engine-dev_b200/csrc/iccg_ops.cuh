// iccg_ops.cuh — one-cell bodies of FdmIccgSolver3's incomplete-Cholesky preconditioner (src/jet/fdm_iccg_solver3.cpp:14-70)
// as __host__ __device__ functions.  The reference evaluates all three recurrences in lexicographic order; a cell only
// depends on its three lower (forward) or upper (backward) neighbours, so every cell of a hyperplane i + j + k = const is
// independent and iccg.cu sweeps hyperplanes — the arithmetic per cell, and therefore the bits, are the reference's.
// tests/hostcheck runs the same source on the CPU against the oracle.
#pragma once
#include "fdm_ops.cuh"

// Preconditioner::build (:14-36): d = 1 / (center - right(i-1)^2 d(i-1) - up(j-1)^2 d(j-1) - front(k-1)^2 d(k-1)), 0 if the
// denominator is 0
template <bool P>
__host__ __device__ __forceinline__ double ic_build_cell(const Mat& A, const double* d, int i, int j, int k, int nx, int ny) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    const double r = (i > 0) ? m_right<P>(A, c - 1) : 0.0, u = (j > 0) ? m_up<P>(A, c - sy) : 0.0, f = (k > 0) ? m_front<P>(A, c - sz) : 0.0;
    const double denom = A.center[c] - ((i > 0) ? (r * r) * d[c - 1] : 0.0) - ((j > 0) ? (u * u) * d[c - sy] : 0.0) -
                         ((k > 0) ? (f * f) * d[c - sz] : 0.0);
    return (fabs(denom) > 0.0) ? 1.0 / denom : 0.0;
}

// Preconditioner::solve, forward substitution (:44-51)
template <bool P>
__host__ __device__ __forceinline__ double ic_forward_cell(const Mat& A, const double* b, const double* y, const double* d, int i,
                                                           int j, int k, int nx, int ny) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    return (b[c] - ((i > 0) ? m_right<P>(A, c - 1) * y[c - 1] : 0.0) - ((j > 0) ? m_up<P>(A, c - sy) * y[c - sy] : 0.0) -
            ((k > 0) ? m_front<P>(A, c - sz) * y[c - sz] : 0.0)) * d[c];
}

// Preconditioner::solve, backward substitution (:53-69)
template <bool P>
__host__ __device__ __forceinline__ double ic_backward_cell(const Mat& A, const double* y, const double* x, const double* d, int i,
                                                            int j, int k, int nx, int ny, int nz) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    return (y[c] - ((i + 1 < nx) ? m_right<P>(A, c) * x[c + 1] : 0.0) - ((j + 1 < ny) ? m_up<P>(A, c) * x[c + sy] : 0.0) -
            ((k + 1 < nz) ? m_front<P>(A, c) * x[c + sz] : 0.0)) * d[c];
}
