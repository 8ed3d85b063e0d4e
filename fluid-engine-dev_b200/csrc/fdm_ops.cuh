// fdm_ops.cuh — one-cell bodies of the FDM linear-system kernels (FdmBlas3 mvm / residual, red-black SOR in the generic
// and the packed face-code form, MG restrict / correct) as __host__ __device__ functions.  fdm.cu's kernels are thin
// wrappers; tests/hostcheck runs the very same source on the CPU against the oracle, bit for bit.
// Every body replays the reference's floating-point expression order (cited per function); built with -fmad=false.
#pragma once
#include <string.h>

#include "common.cuh"

// read-only loads: the texture path on the device, a plain load on the host
template <typename T>
__host__ __device__ __forceinline__ T jc_ldg(const T* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// Off-diagonal accessors.  Packed form: 0.0 - double(weight) * invH^2 exactly as buildSingleSystem forms it
// (src/jet/grid_fractional_single_phase_pressure_solver3.cpp:146-150: `row.right -= term` on a zeroed row).
template <bool P>
__host__ __device__ __forceinline__ double m_right(const Mat& A, size_t c) {
    if (P) { float w = jc_ldg(A.cr + c); return w != 0.f ? -((double)w * A.sx) : 0.0; }
    return jc_ldg(A.right + c);
}
template <bool P>
__host__ __device__ __forceinline__ double m_up(const Mat& A, size_t c) {
    if (P) { float w = jc_ldg(A.cu + c); return w != 0.f ? -((double)w * A.sy) : 0.0; }
    return jc_ldg(A.up + c);
}
template <bool P>
__host__ __device__ __forceinline__ double m_front(const Mat& A, size_t c) {
    if (P) { float w = jc_ldg(A.cf + c); return w != 0.f ? -((double)w * A.sz) : 0.0; }
    return jc_ldg(A.front + c);
}

// The six off-diagonal products in the reference's order -i,+i,-j,+j,-k,+k
// (FdmBlas3::mvm, src/jet/fdm_linear_system3.cpp:79-96; out-of-range terms are literal 0.0).
struct Offdiag {
    double t[6];
};

template <bool P>
__host__ __device__ __forceinline__ Offdiag offdiag_terms(const Mat& A, const double* __restrict__ v, int i, int j, int k,
                                                          int nx, int ny, int nz, size_t c) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    Offdiag o;
    o.t[0] = (i > 0) ? m_right<P>(A, c - 1) * v[c - 1] : 0.0;
    o.t[1] = (i + 1 < nx) ? m_right<P>(A, c) * v[c + 1] : 0.0;
    o.t[2] = (j > 0) ? m_up<P>(A, c - sy) * v[c - sy] : 0.0;
    o.t[3] = (j + 1 < ny) ? m_up<P>(A, c) * v[c + sy] : 0.0;
    o.t[4] = (k > 0) ? m_front<P>(A, c - sz) * v[c - sz] : 0.0;
    o.t[5] = (k + 1 < nz) ? m_front<P>(A, c) * v[c + sz] : 0.0;
    return o;
}

// FdmBlas3::mvm, src/jet/fdm_linear_system3.cpp:79-96.
template <bool P>
__host__ __device__ __forceinline__ void mvm_cell(const Mat& A, const double* __restrict__ v, double* __restrict__ out, int i, int j,
                                                  int k, int nx, int ny, int nz) {
    const size_t c = ((size_t)k * ny + j) * nx + i;
    const Offdiag o = offdiag_terms<P>(A, v, i, j, k, nx, ny, nz, c);
    out[c] = A.center[c] * v[c] + o.t[0] + o.t[1] + o.t[2] + o.t[3] + o.t[4] + o.t[5];
}

// FdmBlas3::residual, src/jet/fdm_linear_system3.cpp:98-116.
template <bool P>
__host__ __device__ __forceinline__ void residual_cell(const Mat& A, const double* __restrict__ x, const double* __restrict__ b,
                                                       double* __restrict__ out, int i, int j, int k, int nx, int ny, int nz) {
    const size_t c = ((size_t)k * ny + j) * nx + i;
    const Offdiag o = offdiag_terms<P>(A, x, i, j, k, nx, ny, nz, c);
    out[c] = b[c] - A.center[c] * x[c] - o.t[0] - o.t[1] - o.t[2] - o.t[3] - o.t[4] - o.t[5];
}

// FdmGaussSeidelSolver3::relaxRedBlack, src/jet/fdm_gauss_seidel_solver3.cpp:150-221: SOR update of one cell,
// x = (1-w) x + w (b - r)/center with r = the six off-diagonal products summed in the reference's order.
template <bool P>
__host__ __device__ __forceinline__ void relax_cell(const Mat& A, const double* __restrict__ b, double* x, double w, int i, int j,
                                                    int k, int nx, int ny, int nz) {
    const size_t c = ((size_t)k * ny + j) * nx + i;
    const Offdiag o = offdiag_terms<P>(A, x, i, j, k, nx, ny, nz, c);
    const double r = o.t[0] + o.t[1] + o.t[2] + o.t[3] + o.t[4] + o.t[5];
    x[c] = (1.0 - w) * x[c] + w * (b[c] - r) / A.center[c];
}

// packed form with face codes
__host__ __device__ __forceinline__ double face_coef(unsigned bits, double s, const float* __restrict__ wts, size_t c) {
    if (bits == 0) return 0.0;
    if (bits == 1) return -s;
    return -((double)jc_ldg(wts + c) * s);
}

__host__ __device__ __forceinline__ double mask_coef(unsigned bit, double neg_s) {  // bit & 1 ? neg_s : +0.0
#ifdef __CUDA_ARCH__
    return __longlong_as_double(__double_as_longlong(neg_s) & -(long long)(bit & 1u));
#else
    long long v;
    memcpy(&v, &neg_s, 8);
    v &= -(long long)(bit & 1u);
    double r;
    memcpy(&r, &v, 8);
    return r;
#endif
}

__host__ __device__ __forceinline__ void relax_cell_code(const Mat& A, const double* __restrict__ b, double* __restrict__ x, double w,
                                                         int i, int j, int k, int nx, int ny, int nz) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    const size_t c = ((size_t)k * ny + j) * nx + i;
    const unsigned own = jc_ldg(A.code + c);
    const unsigned lft = i > 0 ? jc_ldg(A.code + c - 1) : 0u;
    const unsigned dwn = j > 0 ? jc_ldg(A.code + c - sy) : 0u;
    const unsigned bck = k > 0 ? jc_ldg(A.code + c - sz) : 0u;
    // the six 2-bit face classes in the sum order -i,+i,-j,+j,-k,+k (the +face codes are 0 on the domain boundary)
    const unsigned six = (lft & 3u) | ((own & 3u) << 2) | (((dwn >> 2) & 3u) << 4) | (((own >> 2) & 3u) << 6) |
                         (((bck >> 4) & 3u) << 8) | (((own >> 4) & 3u) << 10);
    const double xl = i > 0 ? x[c - 1] : 0.0, xr = i + 1 < nx ? x[c + 1] : 0.0;
    const double xd = j > 0 ? x[c - sy] : 0.0, xu = j + 1 < ny ? x[c + sy] : 0.0;
    const double xb = k > 0 ? x[c - sz] : 0.0, xf = k + 1 < nz ? x[c + sz] : 0.0;
    double cl, cr, cd, cu, cb, cf, center;
    if (six == 0x555u) {  // all six faces fully open liquid-liquid couplings: the bulk of the liquid
        cl = cr = -A.sx; cd = cu = -A.sy; cb = cf = -A.sz;
        // the diagonal of such a cell is a per-level constant (Mat::c6): not read, so that in the bulk whole sectors of
        // the center plane (half of which belong to the other colour anyway) stay in DRAM
        center = A.c6;
    } else if ((six & 0xaaau) == 0) {  // open / closed faces only
        cl = mask_coef(six, -A.sx); cr = mask_coef(six >> 2, -A.sx);
        cd = mask_coef(six >> 4, -A.sy); cu = mask_coef(six >> 6, -A.sy);
        cb = mask_coef(six >> 8, -A.sz); cf = mask_coef(six >> 10, -A.sz);
    } else {  // fractional weights (next to colliders)
        cl = face_coef(six & 3u, A.sx, A.cr, c - 1); cr = face_coef((six >> 2) & 3u, A.sx, A.cr, c);
        cd = face_coef((six >> 4) & 3u, A.sy, A.cu, c - sy); cu = face_coef((six >> 6) & 3u, A.sy, A.cu, c);
        cb = face_coef((six >> 8) & 3u, A.sz, A.cf, c - sz); cf = face_coef((six >> 10) & 3u, A.sz, A.cf, c);
    }
    if (six != 0x555u) center = jc_ldg(A.center + c);
    // absent neighbours carry coefficient 0.0 and value 0.0: the literal 0.0 terms of the reference
    const double r = cl * xl + cr * xr + cd * xd + cu * xu + cb * xb + cf * xf;
    x[c] = (1.0 - w) * x[c] + w * (b[c] - r) / center;
}

// The SOR update split into "gather the cell's coefficients" and "combine them with the seven x values", for kernels that hold
// x somewhere else than in the global array (k_relax_lex_tiles keeps its tile of x in shared memory).  Absent neighbours carry
// coefficient 0.0 and must be given the value 0.0: 0.0 * 0.0 = +0.0 is the literal 0.0 term of the reference, so the sum r — and
// with it the update — has the bits of relax_cell / relax_cell_code (the same products, added in the same order).
struct RelaxCoefs {
    double c[6];     // -i, +i, -j, +j, -k, +k
    double center;
};

template <bool P>
__host__ __device__ __forceinline__ RelaxCoefs relax_coefs(const Mat& A, int i, int j, int k, int nx, int ny, int nz) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    RelaxCoefs q;
    q.c[0] = (i > 0) ? m_right<P>(A, c - 1) : 0.0;
    q.c[1] = (i + 1 < nx) ? m_right<P>(A, c) : 0.0;
    q.c[2] = (j > 0) ? m_up<P>(A, c - sy) : 0.0;
    q.c[3] = (j + 1 < ny) ? m_up<P>(A, c) : 0.0;
    q.c[4] = (k > 0) ? m_front<P>(A, c - sz) : 0.0;
    q.c[5] = (k + 1 < nz) ? m_front<P>(A, c) : 0.0;
    q.center = A.center[c];
    return q;
}

// the face-code form: the coefficients relax_cell_code uses (bulk-diagonal shortcut included).  `six` = the six 2-bit face
// classes of cell c in the sum order -i,+i,-j,+j,-k,+k, from the cell's own code byte and those of its three lower neighbours.
__host__ __device__ __forceinline__ unsigned relax_six(unsigned own, unsigned lft, unsigned dwn, unsigned bck) {
    return (lft & 3u) | ((own & 3u) << 2) | (((dwn >> 2) & 3u) << 4) | (((own >> 2) & 3u) << 6) | (((bck >> 4) & 3u) << 8) |
           (((own >> 4) & 3u) << 10);
}
__host__ __device__ __forceinline__ RelaxCoefs relax_coefs_of_six(const Mat& A, unsigned six, size_t c, size_t sy, size_t sz) {
    RelaxCoefs q;
    if (six == 0x555u) {
        q.c[0] = q.c[1] = -A.sx; q.c[2] = q.c[3] = -A.sy; q.c[4] = q.c[5] = -A.sz;
        q.center = A.c6;
    } else if ((six & 0xaaau) == 0) {
        q.c[0] = mask_coef(six, -A.sx); q.c[1] = mask_coef(six >> 2, -A.sx);
        q.c[2] = mask_coef(six >> 4, -A.sy); q.c[3] = mask_coef(six >> 6, -A.sy);
        q.c[4] = mask_coef(six >> 8, -A.sz); q.c[5] = mask_coef(six >> 10, -A.sz);
    } else {
        q.c[0] = face_coef(six & 3u, A.sx, A.cr, c - 1); q.c[1] = face_coef((six >> 2) & 3u, A.sx, A.cr, c);
        q.c[2] = face_coef((six >> 4) & 3u, A.sy, A.cu, c - sy); q.c[3] = face_coef((six >> 6) & 3u, A.sy, A.cu, c);
        q.c[4] = face_coef((six >> 8) & 3u, A.sz, A.cf, c - sz); q.c[5] = face_coef((six >> 10) & 3u, A.sz, A.cf, c);
    }
    if (six != 0x555u) q.center = jc_ldg(A.center + c);
    return q;
}
__host__ __device__ __forceinline__ RelaxCoefs relax_coefs_code(const Mat& A, int i, int j, int k, int nx, int ny, int nz) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    (void)nz;
    const unsigned own = jc_ldg(A.code + c);
    const unsigned lft = i > 0 ? jc_ldg(A.code + c - 1) : 0u;
    const unsigned dwn = j > 0 ? jc_ldg(A.code + c - sy) : 0u;
    const unsigned bck = k > 0 ? jc_ldg(A.code + c - sz) : 0u;
    return relax_coefs_of_six(A, relax_six(own, lft, dwn, bck), c, sy, sz);
}

// x_new = (1 - w) x + w (b - r) / center with r = the six products in the reference's order
__host__ __device__ __forceinline__ double relax_update(const RelaxCoefs& q, double b, double w, double xc, double xl, double xr,
                                                        double xd, double xu, double xb, double xf) {
    const double r = q.c[0] * xl + q.c[1] * xr + q.c[2] * xd + q.c[3] * xu + q.c[4] * xb + q.c[5] * xf;
    return (1.0 - w) * xc + w * (b - r) / q.center;
}

// the same update on the global array (what the shared-memory tile kernel computes, with absent neighbours read as 0.0):
// tests/hostcheck replays k_relax_lex_tiles with this body
template <bool P, bool CODE>
__host__ __device__ __forceinline__ void relax_cell_split(const Mat& A, const double* __restrict__ b, double* x, double w, int i, int j,
                                                          int k, int nx, int ny, int nz) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny, c = ((size_t)k * ny + j) * nx + i;
    const RelaxCoefs q = CODE ? relax_coefs_code(A, i, j, k, nx, ny, nz) : relax_coefs<P>(A, i, j, k, nx, ny, nz);
    x[c] = relax_update(q, b[c], w, x[c], i > 0 ? x[c - 1] : 0.0, i + 1 < nx ? x[c + 1] : 0.0, j > 0 ? x[c - sy] : 0.0,
                        j + 1 < ny ? x[c + sy] : 0.0, k > 0 ? x[c - sz] : 0.0, k + 1 < nz ? x[c + sz] : 0.0);
}

// FdmMgUtils3::restrict, src/jet/fdm_mg_linear_system3.cpp:43-98: 4x4x4 (1,3,3,1)/8 tensor filter with edge clamp;
// accumulation z-outer, y, x-inner; w = (kx*ky)*kz.  (i, j, k) = COARSE cell.
__host__ __device__ __forceinline__ void restrict_cell(const double* __restrict__ finer, double* __restrict__ coarser, int i, int j,
                                                       int k, int cnx, int cny, int cnz) {
    const double kern[4] = {0.125, 0.375, 0.375, 0.125};
    int iI[4] = {(i > 0) ? 2 * i - 1 : 2 * i, 2 * i, 2 * i + 1, (i + 1 < cnx) ? 2 * i + 2 : 2 * i + 1};
    int jI[4] = {(j > 0) ? 2 * j - 1 : 2 * j, 2 * j, 2 * j + 1, (j + 1 < cny) ? 2 * j + 2 : 2 * j + 1};
    int kI[4] = {(k > 0) ? 2 * k - 1 : 2 * k, 2 * k, 2 * k + 1, (k + 1 < cnz) ? 2 * k + 2 : 2 * k + 1};
    const size_t fnx = 2 * (size_t)cnx, fny = 2 * (size_t)cny;
    double sum = 0.0;
#pragma unroll
    for (int z = 0; z < 4; ++z)
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                double w = kern[x] * kern[y] * kern[z];
                sum += w * finer[((size_t)kI[z] * fny + jI[y]) * fnx + iI[x]];
            }
    coarser[((size_t)k * cny + j) * cnx + i] = sum;
}

// FdmMgUtils3::correct, src/jet/fdm_mg_linear_system3.cpp:100-179: finer += sum_{z,y,x} ((wx*wy)*wz)*coarse.
// The k-clamp uses nz (the reference tests n.y, :159 — identical for ny == nz, out of bounds otherwise).  (i, j, k) = FINE cell.
__host__ __device__ __forceinline__ void correct_cell(const double* __restrict__ coarser, double* __restrict__ finer, int i, int j,
                                                      int k, int nx, int ny, int nz) {
    int iI[2], jI[2], kI[2];
    double iW[2], jW[2], kW[2];
    int ci = i / 2, cj = j / 2, ck = k / 2;
    if ((i & 1) == 0) { iI[0] = (i > 1) ? ci - 1 : ci; iI[1] = ci; iW[0] = 0.25; iW[1] = 0.75; }
    else { iI[0] = ci; iI[1] = (i + 1 < nx) ? ci + 1 : ci; iW[0] = 0.75; iW[1] = 0.25; }
    if ((j & 1) == 0) { jI[0] = (j > 1) ? cj - 1 : cj; jI[1] = cj; jW[0] = 0.25; jW[1] = 0.75; }
    else { jI[0] = cj; jI[1] = (j + 1 < ny) ? cj + 1 : cj; jW[0] = 0.75; jW[1] = 0.25; }
    if ((k & 1) == 0) { kI[0] = (k > 1) ? ck - 1 : ck; kI[1] = ck; kW[0] = 0.25; kW[1] = 0.75; }
    else { kI[0] = ck; kI[1] = (k + 1 < nz) ? ck + 1 : ck; kW[0] = 0.75; kW[1] = 0.25; }
    const size_t cnx = (size_t)nx / 2, cny = (size_t)ny / 2;
    size_t f = ((size_t)k * ny + j) * nx + i;
    double acc = finer[f];
#pragma unroll
    for (int z = 0; z < 2; ++z)
#pragma unroll
        for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                double w = iW[x] * jW[y] * kW[z] * coarser[((size_t)kI[z] * cny + jI[y]) * cnx + iI[x]];
                acc += w;
            }
    finer[f] = acc;
}
