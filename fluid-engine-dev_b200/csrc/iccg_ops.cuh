// iccg_ops.cuh — one-cell bodies of FdmIccgSolver3's incomplete-Cholesky preconditioner (src/jet/fdm_iccg_solver3.cpp:14-70)
// as __host__ __device__ functions.  The reference evaluates all three recurrences in lexicographic order; a cell only
// depends on its three lower (forward) or upper (backward) neighbours, so every cell of a hyperplane i + j + k = const is
// independent and iccg.cu sweeps hyperplanes — the arithmetic per cell, and therefore the bits, are the reference's.
// tests/hostcheck runs the same source on the CPU against the oracle.
#pragma once
#include "fdm_ops.cuh"

// The three recurrences split into "coefficients of the cell" and "combine with the three neighbour values", so that iccg.cu can
// keep a tile of the unknown in shared memory.  Absent neighbours carry coefficient 0.0 and value 0.0: (0.0 * 0.0) * 0.0 and
// 0.0 * 0.0 are +0.0, the literal 0.0 terms of the reference, so the results have the reference's bits.
struct IcCoefs {
    double r, u, f;   // couplings to the three lower (build, forward) or upper (backward) neighbours; 0.0 where absent
};

template <bool P>
__host__ __device__ __forceinline__ IcCoefs ic_lower_coefs(const Mat& A, int i, int j, int k, int nx, int ny) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    IcCoefs q;
    q.r = (i > 0) ? m_right<P>(A, c - 1) : 0.0; q.u = (j > 0) ? m_up<P>(A, c - sy) : 0.0; q.f = (k > 0) ? m_front<P>(A, c - sz) : 0.0;
    return q;
}
template <bool P>
__host__ __device__ __forceinline__ IcCoefs ic_upper_coefs(const Mat& A, int i, int j, int k, int nx, int ny, int nz) {
    const size_t c = ((size_t)k * ny + j) * nx + i;
    IcCoefs q;
    q.r = (i + 1 < nx) ? m_right<P>(A, c) : 0.0; q.u = (j + 1 < ny) ? m_up<P>(A, c) : 0.0; q.f = (k + 1 < nz) ? m_front<P>(A, c) : 0.0;
    return q;
}
// Preconditioner::build (:14-36)
__host__ __device__ __forceinline__ double ic_build_update(const IcCoefs& q, double center, double dl, double dd, double db) {
    const double denom = center - (q.r * q.r) * dl - (q.u * q.u) * dd - (q.f * q.f) * db;
    return (fabs(denom) > 0.0) ? 1.0 / denom : 0.0;
}
// Preconditioner::solve, forward (:44-51) and backward (:53-69) substitution: (rhs - r v_x - u v_y - f v_z) * d
__host__ __device__ __forceinline__ double ic_sub_update(const IcCoefs& q, double rhs, double vx, double vy, double vz, double d) {
    return (rhs - q.r * vx - q.u * vy - q.f * vz) * d;
}

// Preconditioner::build (:14-36): d = 1 / (center - right(i-1)^2 d(i-1) - up(j-1)^2 d(j-1) - front(k-1)^2 d(k-1)), 0 if the
// denominator is 0
template <bool P>
__host__ __device__ __forceinline__ double ic_build_cell(const Mat& A, const double* d, int i, int j, int k, int nx, int ny) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    return ic_build_update(ic_lower_coefs<P>(A, i, j, k, nx, ny), A.center[c], (i > 0) ? d[c - 1] : 0.0, (j > 0) ? d[c - sy] : 0.0,
                           (k > 0) ? d[c - sz] : 0.0);
}

// Preconditioner::solve, forward substitution (:44-51)
template <bool P>
__host__ __device__ __forceinline__ double ic_forward_cell(const Mat& A, const double* b, const double* y, const double* d, int i,
                                                           int j, int k, int nx, int ny) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    return ic_sub_update(ic_lower_coefs<P>(A, i, j, k, nx, ny), b[c], (i > 0) ? y[c - 1] : 0.0, (j > 0) ? y[c - sy] : 0.0,
                         (k > 0) ? y[c - sz] : 0.0, d[c]);
}

// Preconditioner::solve, backward substitution (:53-69)
template <bool P>
__host__ __device__ __forceinline__ double ic_backward_cell(const Mat& A, const double* y, const double* x, const double* d, int i,
                                                            int j, int k, int nx, int ny, int nz) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    return ic_sub_update(ic_upper_coefs<P>(A, i, j, k, nx, ny, nz), y[c], (i + 1 < nx) ? x[c + 1] : 0.0, (j + 1 < ny) ? x[c + sy] : 0.0,
                         (k + 1 < nz) ? x[c + sz] : 0.0, d[c]);
}
