// sampling.cuh — device replays of Jet's grid samplers and level-set fraction helpers.
// Expression order follows the reference exactly (files cited per function); built with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

// View of one scalar array with Jet's sampler parameters (spacing + data origin).
struct GridView {
    const double* data;
    int nx, ny, nz;
    double hx, hy, hz;
    double ox, oy, oz;
    double ihx = 0.0, ihy = 0.0, ihz = 0.0;  // 1/h where h is an exact power of two, else 0 (see jc_norm)
};

// (x - o) / h as the reference's samplers normalise positions (array_samplers3-inl.h:123,206).  An IEEE division by a
// power of two is an exact scaling, so it equals the multiplication by the (exact) reciprocal bit for bit; the host
// passes ih = 1/h only when h is a power of two (ih == 0 selects the literal division).  FP64 division costs ~30
// instructions and the advection kernel needs 90 of them per particle, so this matters on 2^-k grids.
__host__ __device__ __forceinline__ double jc_norm(double x, double o, double h, double ih) {
    return ih != 0.0 ? (x - o) * ih : (x - o) / h;
}
__host__ __device__ __forceinline__ float jc_normf(float x, float o, float h, float ih) {
    return ih != 0.0f ? (x - o) * ih : (x - o) / h;
}

// Host: 1/h if h is a power of two (then the reciprocal is exact), else 0.
static inline double jc_pow2_inverse(double h) {
    int e = 0;
    return (h > 0.0 && frexp(h, &e) == 0.5) ? 1.0 / h : 0.0;
}

// getBarycentric, include/jet/detail/math_utils-inl.h:146-174 (iLow == 0 on this path).
__host__ __device__ __forceinline__ void jc_barycentric(double x, int iHigh, int* i, double* f) {
    // s = floor(x) is integer-valued, so the reference's integer tests (i < 0, i > iHigh - 1 on a ssize_t) are decided
    // on the double itself — the same branches for every finite x, without a 64-bit float-to-int conversion; wild
    // inputs (|x| beyond the int range) land in the clamps instead of overflowing the conversion.
    const double s = floor(x);
    if (iHigh == 0) { *i = 0; *f = 0.0; }
    else if (s < 0.0) { *i = 0; *f = 0.0; }
    else if (s > (double)(iHigh - 1)) { *i = iHigh - 1; *f = 1.0; }
    else { *i = (int)s; *f = x - s; }
}

__host__ __device__ __forceinline__ void jc_barycentricf(float x, int iHigh, int* i, float* f) {
    const float s = floorf(x);
    if (iHigh == 0) { *i = 0; *f = 0.0f; }
    else if (s < 0.0f) { *i = 0; *f = 0.0f; }
    else if (s > (float)(iHigh - 1)) { *i = iHigh - 1; *f = 1.0f; }
    else { *i = (int)s; *f = x - s; }
}

__host__ __device__ __forceinline__ double jc_lerp(double a, double b, double f) { return (1 - f) * a + f * b; }
__host__ __device__ __forceinline__ float jc_lerpf(float a, float b, float f) { return (1 - f) * a + f * b; }

// LinearArraySampler3<double,double>::operator(), include/jet/detail/array_samplers3-inl.h:115-149:
// normalise with a DIVISION by h; nested lerp x -> y -> z (math_utils-inl.h:176-211).
__host__ __device__ __forceinline__ double jc_sample(const GridView& g, double px, double py, double pz) {
    int i, j, k;
    double fx, fy, fz;
    jc_barycentric(jc_norm(px, g.ox, g.hx, g.ihx), g.nx - 1, &i, &fx);
    jc_barycentric(jc_norm(py, g.oy, g.hy, g.ihy), g.ny - 1, &j, &fy);
    jc_barycentric(jc_norm(pz, g.oz, g.hz, g.ihz), g.nz - 1, &k, &fz);
    int ip1 = min(i + 1, g.nx - 1), jp1 = min(j + 1, g.ny - 1), kp1 = min(k + 1, g.nz - 1);
    const size_t sy = (size_t)g.nx, sz = (size_t)g.nx * g.ny;
    const double* d = g.data;
    size_t r00 = (size_t)k * sz + (size_t)j * sy, r10 = (size_t)k * sz + (size_t)jp1 * sy;
    size_t r01 = (size_t)kp1 * sz + (size_t)j * sy, r11 = (size_t)kp1 * sz + (size_t)jp1 * sy;
    double a = jc_lerp(jc_lerp(d[r00 + i], d[r00 + ip1], fx), jc_lerp(d[r10 + i], d[r10 + ip1], fx), fy);
    double b = jc_lerp(jc_lerp(d[r01 + i], d[r01 + ip1], fx), jc_lerp(d[r11 + i], d[r11 + ip1], fx), fy);
    return jc_lerp(a, b, fz);
}

// LinearArraySampler3<float,float>::operator() as FlipSolver3's delta sampler uses it (src/jet/flip_solver3.cpp:72-88).
__host__ __device__ __forceinline__ float jc_samplef(const float* d, int nx, int ny, int nz, float hx, float hy, float hz,
                                            float ox, float oy, float oz, float px, float py, float pz,
                                            float ihx = 0.f, float ihy = 0.f, float ihz = 0.f) {
    int i, j, k;
    float fx, fy, fz;
    jc_barycentricf(jc_normf(px, ox, hx, ihx), nx - 1, &i, &fx);
    jc_barycentricf(jc_normf(py, oy, hy, ihy), ny - 1, &j, &fy);
    jc_barycentricf(jc_normf(pz, oz, hz, ihz), nz - 1, &k, &fz);
    int ip1 = min(i + 1, nx - 1), jp1 = min(j + 1, ny - 1), kp1 = min(k + 1, nz - 1);
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    size_t r00 = (size_t)k * sz + (size_t)j * sy, r10 = (size_t)k * sz + (size_t)jp1 * sy;
    size_t r01 = (size_t)kp1 * sz + (size_t)j * sy, r11 = (size_t)kp1 * sz + (size_t)jp1 * sy;
    float a = jc_lerpf(jc_lerpf(d[r00 + i], d[r00 + ip1], fx), jc_lerpf(d[r10 + i], d[r10 + ip1], fx), fy);
    float b = jc_lerpf(jc_lerpf(d[r01 + i], d[r01 + ip1], fx), jc_lerpf(d[r11 + i], d[r11 + ip1], fx), fy);
    return jc_lerpf(a, b, fz);
}

// Base index + fractions of LinearArraySampler3::getCoordinatesAndWeights (array_samplers3-inl.h:151-193):
// normalises with a MULTIPLICATION by invH = 1/h.
struct BaseFrac {
    int i, j, k, ip1, jp1, kp1;
    double fx, fy, fz;
};

__host__ __device__ __forceinline__ BaseFrac jc_base_mul(int nx, int ny, int nz, double ihx, double ihy, double ihz, double ox,
                                                double oy, double oz, double px, double py, double pz) {
    BaseFrac b;
    jc_barycentric((px - ox) * ihx, nx - 1, &b.i, &b.fx);
    jc_barycentric((py - oy) * ihy, ny - 1, &b.j, &b.fy);
    jc_barycentric((pz - oz) * ihz, nz - 1, &b.k, &b.fz);
    b.ip1 = min(b.i + 1, nx - 1); b.jp1 = min(b.j + 1, ny - 1); b.kp1 = min(b.k + 1, nz - 1);
    return b;
}

// Same with the division form used by operator() and getCoordinatesAndGradientWeights (:195-261).
__host__ __device__ __forceinline__ BaseFrac jc_base_div(int nx, int ny, int nz, double hx, double hy, double hz, double ox,
                                                double oy, double oz, double px, double py, double pz,
                                                double ihx = 0.0, double ihy = 0.0, double ihz = 0.0) {
    BaseFrac b;
    jc_barycentric(jc_norm(px, ox, hx, ihx), nx - 1, &b.i, &b.fx);
    jc_barycentric(jc_norm(py, oy, hy, ihy), ny - 1, &b.j, &b.fy);
    jc_barycentric(jc_norm(pz, oz, hz, ihz), nz - 1, &b.k, &b.fz);
    b.ip1 = min(b.i + 1, nx - 1); b.jp1 = min(b.j + 1, ny - 1); b.kp1 = min(b.k + 1, nz - 1);
    return b;
}

// Corner q (0..7) weight in the reference's corner order (i,j,k),(i+1,j,k),(i,j+1,k),(i+1,j+1,k),(i,j,k+1)...
__host__ __device__ __forceinline__ double jc_corner_weight(const BaseFrac& b, int q) {
    double wx = (q & 1) ? b.fx : (1 - b.fx);
    double wy = (q & 2) ? b.fy : (1 - b.fy);
    double wz = (q & 4) ? b.fz : (1 - b.fz);
    return wx * wy * wz;
}

// fractionInsideSdf, include/jet/detail/level_set_utils-inl.h:52-63.
__host__ __device__ __forceinline__ double jc_frac_inside_sdf(double p0, double p1) {
    if (p0 < 0 && p1 < 0) return 1;
    if (p0 < 0 && !(p1 < 0)) return p0 / (p0 - p1);
    if (!(p0 < 0) && p1 < 0) return p1 / (p1 - p0);
    return 0;
}

// fractionInside (marching-squares covered area), level_set_utils-inl.h:72-136.
__host__ __device__ __forceinline__ double jc_frac_inside4(double bl, double br, double tl, double tr) {
    int cnt = (bl < 0) + (tl < 0) + (br < 0) + (tr < 0);
    double l0 = bl, l1 = br, l2 = tr, l3 = tl;
#define JC_CYCLE() { double t = l0; l0 = l1; l1 = l2; l2 = l3; l3 = t; }
    if (cnt == 4) return 1;
    if (cnt == 3) {
        for (int g = 0; g < 4 && l0 < 0; ++g) JC_CYCLE();
        double s0 = 1 - jc_frac_inside_sdf(l0, l3);
        double s1 = 1 - jc_frac_inside_sdf(l0, l1);
        return 1 - 0.5 * s0 * s1;
    }
    if (cnt == 2) {
        for (int g = 0; g < 4 && (l0 >= 0 || !(l1 < 0 || l2 < 0)); ++g) JC_CYCLE();
        if (l1 < 0) {
            double sl = jc_frac_inside_sdf(l0, l3);
            double sr = jc_frac_inside_sdf(l1, l2);
            return 0.5 * (sl + sr);
        }
        double mid = 0.25 * (l0 + l1 + l2 + l3);
        if (mid < 0) {
            double area = 0;
            double s1 = 1 - jc_frac_inside_sdf(l0, l3);
            double s3 = 1 - jc_frac_inside_sdf(l2, l3);
            area += 0.5 * s1 * s3;
            double s2 = 1 - jc_frac_inside_sdf(l2, l1);
            double s0 = 1 - jc_frac_inside_sdf(l0, l1);
            area += 0.5 * s0 * s2;
            return 1 - area;
        } else {
            double area = 0;
            double s0 = jc_frac_inside_sdf(l0, l1);
            double s1 = jc_frac_inside_sdf(l0, l3);
            area += 0.5 * s0 * s1;
            double s2 = jc_frac_inside_sdf(l2, l1);
            double s3 = jc_frac_inside_sdf(l2, l3);
            area += 0.5 * s2 * s3;
            return area;
        }
    }
    if (cnt == 1) {
        for (int g = 0; g < 4 && l0 >= 0; ++g) JC_CYCLE();
        double s0 = jc_frac_inside_sdf(l0, l3);
        double s1 = jc_frac_inside_sdf(l0, l1);
        return 0.5 * s0 * s1;
    }
#undef JC_CYCLE
    return 0;
}

__host__ __device__ __forceinline__ double jc_clamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
