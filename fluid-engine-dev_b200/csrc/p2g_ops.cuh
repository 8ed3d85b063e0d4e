// p2g_ops.cuh — arithmetic of the particle-to-grid transfer as __host__ __device__ functions: 128-bit fixed-point
// conversion and accumulation, the per-cell register accumulation, the tile flush and the normalisation.  hybrid.cu's
// kernels are wrappers (shared memory, thread loops, __syncthreads around them); tests/hostcheck runs the same source on
// the CPU — atomics become plain additions there — against exact integer arithmetic and the oracle.
#pragma once
#include <math.h>
#include "particle_ops.cuh"

#define NATTR 15  // px,py,pz, vx,vy,vz, cX(3), cY(3), cZ(3)

// Owning cell of a particle (our own binning for locality, NOT the reference's hash grid): clamped floor.
__host__ __device__ __forceinline__ int cell_coord(double x, double o, double ih, int n) {
    double s = floor((x - o) * ih);
    int c = (s < 0.0) ? 0 : ((s > (double)(n - 1)) ? n - 1 : (int)s);
    return c;
}


struct AttrPtrs {
    double* p[NATTR];
};


#define TC 8
#define P2G_THREADS 256
#define P2G_KW 50          // weight sums: fixed point with 50 fractional bits (sum of weights per face < 4096)
#define P2G_HEADROOM 12    // value sums: max|value| * 2^HEADROOM * 2^k < 2^62

struct P2GScale {
    double scale, inv_scale;  // 2^k, 2^-k
};


// atomics of the accumulation: real atomics on the device, plain read-modify-write in the single-threaded host check
template <typename T>
__host__ __device__ __forceinline__ T jc_atomic_add(T* p, T v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    const T old = *p;
    *p = old + v;
    return old;
#endif
}
__host__ __device__ __forceinline__ unsigned long long jc_double2ull_rd(double x) {  // x >= 0: round down == truncate
#ifdef __CUDA_ARCH__
    return __double2ull_rd(x);
#else
    return (unsigned long long)x;
#endif
}

// fixed-point scale of one step from the largest value a particle can deposit
__host__ __device__ __forceinline__ P2GScale p2g_scale_from_vmax(double vmax) {
    int e = (vmax > 0.0 && isfinite(vmax)) ? ilogb(vmax) + 1 : 0;  // vmax < 2^e
    int k = 62 - P2G_HEADROOM - e;
    P2GScale s;
    s.scale = ldexp(1.0, k);
    s.inv_scale = ldexp(1.0, -k);
    return s;
}

struct P2GAcc {  // global accumulators; the high words of the value sums live in the velocity arrays themselves
    unsigned long long* sum[3];     // value sum, high word
    unsigned long long* sum_lo[3];  // value sum, low word
    unsigned long long* wsum[3];    // weight sum, high word
    unsigned long long* wsum_lo[3]; // weight sum, low word
    char* marker[3];
};

// |t| < 2^63 -> signed 128-bit fixed point with 64 fractional bits (two's complement in hi:lo)
__host__ __device__ __forceinline__ void to_fixed128(double t, unsigned long long* hi, unsigned long long* lo) {
    const double a = fabs(t), fl = floor(a);
    unsigned long long h = (unsigned long long)fl;
    unsigned long long l = jc_double2ull_rd((a - fl) * 18446744073709551616.0);  // exact difference, < 2^64
    if (t < 0.0) { l = 0ull - l; h = ~h + (l == 0ull ? 1ull : 0ull); }
    *hi = h; *lo = l;
}

// (HI:LO) += (hi:lo) with two 64-bit atomics; the number of carries out of LO is fixed by the final sum, so the
// result is independent of the order in which contributions arrive.
__host__ __device__ __forceinline__ void acc128(unsigned long long* HI, unsigned long long* LO, unsigned long long hi,
                                       unsigned long long lo) {
    if (lo) {
        const unsigned long long old = jc_atomic_add(LO, lo);
        hi += (old > ~lo) ? 1ull : 0ull;
    }
    if (hi) jc_atomic_add(HI, hi);
}

// The same on a shared-memory accumulator stored as four 32-bit limbs `stride` words apart (non-negative addend).
__host__ __device__ __forceinline__ void acc_limbs(unsigned* L, int stride, unsigned long long hi, unsigned long long lo) {
    const unsigned x[4] = {(unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32)};
    unsigned carry = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long add = (unsigned long long)x[j] + carry;
        const unsigned a32 = (unsigned)add;
        carry = (unsigned)(add >> 32);
        if (a32) {
            const unsigned old = jc_atomic_add(L + j * stride, a32);
            carry += (old > ~a32) ? 1u : 0u;
        }
    }
}

__host__ __device__ __forceinline__ double from_fixed128(unsigned long long hi, unsigned long long lo) {  // value * 2^64
    const bool neg = (long long)hi < 0;
    if (neg) { lo = 0ull - lo; hi = ~hi + (lo == 0ull ? 1ull : 0ull); }
    const double v = (double)hi * 18446744073709551616.0 + (double)lo;
    return neg ? -v : v;
}


// ---- P2G, per-cell form ------------------------------------------------------------------------------------------
// All particles of one cell deposit into the same small window of faces: 2 along the component's own (node-aligned)
// axis x 3 x 3 along the two cell-centred axes = 18 faces.  One thread owns a cell, sums its particles' contributions
// to those 18 faces in REGISTERS (FP64, in the cell's fixed order: ascending original index), and only then converts
// the 18 partial sums to fixed point and adds them to the tile's shared accumulators: 18 conversions + limb atomics per
// cell and component instead of 8 per particle (x8 particles).  The cross-cell / cross-tile / cross-rank combination is
// still integer addition, so the result is independent of tile order and of the slab decomposition (a ghost cell is
// summed by the very same expression on every rank that holds it).
// Slots outside a particle's actual 2x2x2 corners get weight 0 (val = +-0): adding them changes nothing.
__host__ __device__ constexpr unsigned p2g_combo_mask(int ox, int oy, int oz) {
    unsigned m = 0;
    for (int c = oz; c <= oz + 1; ++c)
        for (int b = oy; b <= oy + 1; ++b)
            for (int a = ox; a <= ox + 1; ++a) m |= 1u << ((c * 3 + b) * 3 + a);
    return m;
}

// Carry-free shared accumulator (experimental, JETCU_P2G_KERNEL=carryfree): the 128-bit non-negative integer is spread
// over five 32-bit words holding 26 bits each, so an addition is five independent fire-and-forget adds (RED) with no
// carry to propagate — acc_limbs chains up to four atomics through their return values.  A word overflows only after 64
// additions; a face receives at most 18 per tile and component here.
__host__ __device__ __forceinline__ void acc_limbs26(unsigned* L, int stride, unsigned long long hi, unsigned long long lo) {
    const unsigned M = (1u << 26) - 1u;
    const unsigned l0 = (unsigned)lo & M, l1 = (unsigned)(lo >> 26) & M;
    const unsigned l2 = (unsigned)((lo >> 52) | (hi << 12)) & M, l3 = (unsigned)(hi >> 14) & M, l4 = (unsigned)(hi >> 40);
    if (l0) jc_atomic_add(L, l0);
    if (l1) jc_atomic_add(L + stride, l1);
    if (l2) jc_atomic_add(L + 2 * stride, l2);
    if (l3) jc_atomic_add(L + 3 * stride, l3);
    if (l4) jc_atomic_add(L + 4 * stride, l4);
}

__host__ __device__ __forceinline__ void limbs26_to_u128(const unsigned* L, int stride, unsigned long long* hi, unsigned long long* lo) {
    unsigned long long h = 0ull, l = 0ull;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const unsigned long long v = L[j * stride];
        const int sh = 26 * j;
        const unsigned long long plo = sh < 64 ? v << sh : 0ull;
        const unsigned long long phi = sh == 0 ? 0ull : (sh < 64 ? v >> (64 - sh) : v << (sh - 64));
        l += plo;
        h += phi + (l < plo ? 1ull : 0ull);
    }
    *hi = h; *lo = l;
}

template <bool APIC, int AX, bool CF>
__host__ __device__ __forceinline__ void p2g_cell_component(const GridD& g, int ci, int cj, int ck, uint32_t s0, uint32_t e0,
                                                   const AttrPtrs& a, double scale, int I0, int J0, int K0, unsigned* sV,
                                                   unsigned* sW, unsigned char* sM, const P2GAcc& acc) {
    constexpr int NF = (TC + 1) * (TC + 2) * (TC + 2);
    constexpr int NX = AX == 0 ? 2 : 3, NY = AX == 1 ? 2 : 3, NZ = AX == 2 ? 2 : 3;
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    const double ox = g.ox + 0.5 * (AX == 0 ? 0.0 : g.hx), oy = g.oy + 0.5 * (AX == 1 ? 0.0 : g.hy),
                 oz = g.oz + 0.5 * (AX == 2 ? 0.0 : g.hz);
    const double ihx = 1.0 / g.hx, ihy = 1.0 / g.hy, ihz = 1.0 / g.hz;
    const double wscale = 1125899906842624.0;  // 2^50
    // window origin in face indices
    const int bx0 = ci - (AX == 0 ? 0 : 1), by0 = cj - (AX == 1 ? 0 : 1), bz0 = ck - (AX == 2 ? 0 : 1);
    double sv[NZ][NY][NX], sw[NZ][NY][NX];
#pragma unroll
    for (int c = 0; c < NZ; ++c)
#pragma unroll
        for (int b = 0; b < NY; ++b)
#pragma unroll
            for (int x = 0; x < NX; ++x) { sv[c][b][x] = 0.0; sw[c][b][x] = 0.0; }
    unsigned touched = 0;  // bit ((c*3 + b)*3 + a): slot is a real corner of at least one particle (markers)
    for (uint32_t q = s0; q < e0; ++q) {
        double px = a.p[0][q], py = a.p[1][q], pz = a.p[2][q];
        if (APIC) {  // tangential clamp to [lo + h/2, hi - h/2] (apic_solver3.cpp:76-84,94-102,112-120)
            if (AX != 0) px = jc_clamp(px, g.ox + g.hx / 2.0, (g.ox + g.hx * (double)g.nx) - g.hx / 2.0);
            if (AX != 1) py = jc_clamp(py, g.oy + g.hy / 2.0, (g.oy + g.hy * (double)g.ny) - g.hy / 2.0);
            if (AX != 2) pz = jc_clamp(pz, g.oz + g.hz / 2.0, (g.oz + g.hz * (double)g.nz) - g.hz / 2.0);
        }
        const BaseFrac bf = jc_base_mul(sx, sy, sz, ihx, ihy, ihz, ox, oy, oz, px, py, pz);
        const double pv = a.p[3 + AX][q];
        double c0 = 0.0, c1 = 0.0, c2 = 0.0;
        if (APIC) { c0 = a.p[6 + 3 * AX][q]; c1 = a.p[7 + 3 * AX][q]; c2 = a.p[8 + 3 * AX][q]; }
        const int dx = bf.i - bx0, dy = bf.j - by0, dz = bf.k - bz0;  // 0 on the node axis, 0 or 1 on the centred ones
        const bool inwin = (unsigned)dx <= (AX == 0 ? 0u : 1u) && (unsigned)dy <= (AX == 1 ? 0u : 1u) &&
                           (unsigned)dz <= (AX == 2 ? 0u : 1u);
        if (!inwin) {  // cannot happen for a particle binned into this cell; kept so that nothing is ever dropped
#pragma unroll
            for (int cnr = 0; cnr < 8; ++cnr) {
                const int fi = (cnr & 1) ? bf.ip1 : bf.i, fj = (cnr & 2) ? bf.jp1 : bf.j, fk = (cnr & 4) ? bf.kp1 : bf.k;
                const double wgt = jc_corner_weight(bf, cnr);
                double val = pv * wgt;
                if (APIC) {
                    const double gx = ox + g.hx * (double)fi, gy = oy + g.hy * (double)fj, gz = oz + g.hz * (double)fk;
                    val = wgt * (pv + (c0 * (gx - px) + c1 * (gy - py) + c2 * (gz - pz)));
                }
                unsigned long long vh, vl, wh, wl;
                to_fixed128(val * scale, &vh, &vl);
                to_fixed128(wgt * wscale, &wh, &wl);
                const size_t f = ((size_t)fk * sy + fj) * sx + fi;
                acc128(acc.sum[AX] + f, acc.sum_lo[AX] + f, vh, vl);
                acc128(acc.wsum[AX] + f, acc.wsum_lo[AX] + f, wh, wl);
                acc.marker[AX][f] = 1;
            }
            continue;
        }
        // the two real weights per axis placed at offset d, zero elsewhere
        double Wx[3], Wy[3], Wz[3];
        Wx[0] = dx == 0 ? 1 - bf.fx : 0.0; Wx[1] = dx == 0 ? bf.fx : 1 - bf.fx; Wx[2] = dx == 0 ? 0.0 : bf.fx;
        Wy[0] = dy == 0 ? 1 - bf.fy : 0.0; Wy[1] = dy == 0 ? bf.fy : 1 - bf.fy; Wy[2] = dy == 0 ? 0.0 : bf.fy;
        Wz[0] = dz == 0 ? 1 - bf.fz : 0.0; Wz[1] = dz == 0 ? bf.fz : 1 - bf.fz; Wz[2] = dz == 0 ? 0.0 : bf.fz;
#pragma unroll
        for (int c = 0; c < NZ; ++c)
#pragma unroll
            for (int b = 0; b < NY; ++b)
#pragma unroll
                for (int x = 0; x < NX; ++x) {
                    const double wgt = Wx[x] * Wy[b] * Wz[c];  // (wx*wy)*wz as in array_samplers3-inl.h:176-192
                    double val;
                    if (APIC) {
                        const double gx = ox + g.hx * (double)(bx0 + x), gy = oy + g.hy * (double)(by0 + b),
                                     gz = oz + g.hz * (double)(bz0 + c);
                        val = wgt * (pv + (c0 * (gx - px) + c1 * (gy - py) + c2 * (gz - pz)));
                    } else {
                        val = pv * wgt;
                    }
                    sv[c][b][x] += val; sw[c][b][x] += wgt;
                }
        // dx/dy/dz in {0,1}: one of 8 corner-block positions inside the window
        const int combo = dx + 2 * dy + 4 * dz;
        touched |= combo == 0 ? p2g_combo_mask(0, 0, 0) : combo == 1 ? p2g_combo_mask(1, 0, 0) : combo == 2 ? p2g_combo_mask(0, 1, 0)
                 : combo == 3 ? p2g_combo_mask(1, 1, 0) : combo == 4 ? p2g_combo_mask(0, 0, 1) : combo == 5 ? p2g_combo_mask(1, 0, 1)
                 : combo == 6 ? p2g_combo_mask(0, 1, 1) : p2g_combo_mask(1, 1, 1);
    }
    if (!touched) return;
    // tile frame of this component (tile + one ring on the centred axes)
    const int lx0 = I0 - (AX == 0 ? 0 : 1), ly0 = J0 - (AX == 1 ? 0 : 1), lz0 = K0 - (AX == 2 ? 0 : 1);
    constexpr int TDX = TC + (AX == 0 ? 1 : 2), TDY = TC + (AX == 1 ? 1 : 2);
#pragma unroll
    for (int c = 0; c < NZ; ++c)
#pragma unroll
        for (int b = 0; b < NY; ++b)
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                if (!(touched >> ((c * 3 + b) * 3 + x) & 1u)) continue;
                unsigned long long vh, vl, wh, wl;
                to_fixed128(sv[c][b][x] * scale, &vh, &vl);
                to_fixed128(sw[c][b][x] * wscale, &wh, &wl);
                // bias the value by TWICE the weight: |sum v*w| <= max|v| * sum w < 2^e * sum w up to the rounding of the
                // FP64 partial sums, so value + 2 * 2^e * weight is safely non-negative
                const unsigned long long w2l = wl << 1, w2h = (wh << 1) | (wl >> 63);
                const unsigned long long bl = vl + w2l, bh = vh + w2h + (bl < vl ? 1ull : 0ull);
                const int t = ((bz0 + c - lz0) * TDY + (by0 + b - ly0)) * TDX + (bx0 + x - lx0);
                if (CF) { acc_limbs26(sV + t, NF, bh, bl); acc_limbs26(sW + t, NF, wh, wl); }
                else { acc_limbs(sV + t, NF, bh, bl); acc_limbs(sW + t, NF, wh, wl); }
                sM[t] = 1;
            }
}


// One face of a tile: the shared accumulators (value + 2 * weight bias, weight) are rebuilt as 128-bit integers, the bias
// is removed exactly and the face is added to the global accumulators.  t = index inside component ax's tile frame.
template <bool CF>
__host__ __device__ __forceinline__ void p2g_tile_flush_face(const GridD& g, int I0, int J0, int K0, int ax, int t,
                                                             const unsigned* sV, const unsigned* sW, const unsigned char* sM,
                                                             const P2GAcc& acc) {
    constexpr int NF = (TC + 1) * (TC + 2) * (TC + 2);
    if (!sM[t]) return;
    const int sx = g.nx + (ax == 0), sy = g.ny + (ax == 1);
    const int lx0 = I0 - (ax == 0 ? 0 : 1), ly0 = J0 - (ax == 1 ? 0 : 1), lz0 = K0 - (ax == 2 ? 0 : 1);
    const int dx = TC + (ax == 0 ? 1 : 2), dy = TC + (ax == 1 ? 1 : 2);
    const int lx = t % dx, ly = (t / dx) % dy, lz = t / (dx * dy);
    const size_t f = ((size_t)(lz0 + lz) * sy + (ly0 + ly)) * sx + (lx0 + lx);
    unsigned long long wl, wh, vl, vh;
    if (CF) {
        limbs26_to_u128(sW + t, NF, &wh, &wl);
        limbs26_to_u128(sV + t, NF, &vh, &vl);
    } else {
        wl = (unsigned long long)sW[t] | ((unsigned long long)sW[NF + t] << 32);
        wh = (unsigned long long)sW[2 * NF + t] | ((unsigned long long)sW[3 * NF + t] << 32);
        vl = (unsigned long long)sV[t] | ((unsigned long long)sV[NF + t] << 32);
        vh = (unsigned long long)sV[2 * NF + t] | ((unsigned long long)sV[3 * NF + t] << 32);
    }
    const unsigned long long w2l = wl << 1, w2h = (wh << 1) | (wl >> 63);
    vh = vh - w2h - (vl < w2l ? 1ull : 0ull); vl = vl - w2l;  // remove the bias: exact 128-bit subtraction
    acc128(acc.sum[ax] + f, acc.sum_lo[ax] + f, vh, vl);
    acc128(acc.wsum[ax] + f, acc.wsum_lo[ax] + f, wh, wl);
    acc.marker[ax][f] = 1;
}

// u = sum / wsum where wsum > 0 (pic_solver3.cpp:178-192), back in double; FLIP's float snapshot
// (src/jet/flip_solver3.cpp:39-47) is taken in the same pass.  err[0] is raised if a weight sum left the
// fixed-point headroom (more than 2^HEADROOM particles' worth of weight on one face).
__host__ __device__ __forceinline__ void p2g_normalise_face(double* u, const unsigned long long* sum_lo, const unsigned long long* wsum,
                                                            const unsigned long long* wsum_lo, double inv_scale, float* snapshot,
                                                            int* err, size_t f) {
    const double S = from_fixed128(reinterpret_cast<const unsigned long long*>(u)[f], sum_lo[f]);  // value * 2^(k+64)
    const unsigned long long Wh = wsum[f], Wl = wsum_lo[f];
    double val = S * 5.421010862427522e-20 * inv_scale;  // * 2^-64 * 2^-k (exact scalings)
    if (Wh | Wl) {
        const double wd = from_fixed128(Wh, Wl) * 5.421010862427522e-20 * 8.8817841970012523e-16;  // * 2^-64 * 2^-50
        val = val / wd;
        if (wd > (double)(1 << P2G_HEADROOM)) err[0] = 1;
    }
    u[f] = val;
    if (snapshot) snapshot[f] = (float)val;
}
