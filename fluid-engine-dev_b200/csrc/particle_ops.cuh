// particle_ops.cuh — per-particle bodies of the G2P and advection kernels as __host__ __device__ functions: hybrid.cu's
// kernels are thin wrappers (SoA load / store around them), and tests/hostcheck runs the very same source on the CPU
// against the oracle (bit for bit: these stages have no reductions).
#pragma once
#include "collider_ops.cuh"

// POW2: all three spacings are powers of two (GridD::ih* != 0), so (x - o) / h is evaluated as the exact product with
// 1/h and the division path is not even compiled in.  I = index type: int when every component grid has < 2^31
// samples (always, up to ~1280^3), else long long — 64-bit integer multiplies were a third of this kernel's instructions.
template <bool POW2>
__host__ __device__ __forceinline__ double norm_pos(double x, double o, double h, double ih) {
    return POW2 ? (x - o) * ih : jc_norm(x, o, h, ih);
}

// PAIRS: the component arrays are "pair arrays" — entry i holds (d[i], d[min(i + 1, sx - 1)]) as a 16-byte double2
// (k_pack_pairs), so the two x-neighbours of a trilinear sample arrive in ONE 128-bit load.  The advection kernel is bound
// by the number of L1 load wavefronts (ncu: l1tex data-pipe 90 % busy, FP64 43 %), not by DRAM; pairs halve them.  The
// values are the same doubles, so the results are bit-identical to the plain-array form.
template <bool PAIRS>
__host__ __device__ __forceinline__ void jc_load_x2(const double* __restrict__ d, size_t lo, size_t hi, double* a, double* b) {
    if (PAIRS) {
        const double2 q = reinterpret_cast<const double2*>(d)[lo];
        *a = q.x; *b = q.y;
    } else {
        *a = d[lo]; *b = d[hi];
    }
}

template <bool PAIRS, typename I>
__host__ __device__ __forceinline__ double sample_rows(const double* __restrict__ d, I r00, I r10, I r01, I r11, int i, int i1, double fx,
                                              double fy, double fz) {
    double a0, a1, b0, b1, c0, c1, d0, d1;
    jc_load_x2<PAIRS>(d, (size_t)(r00 + i), (size_t)(r00 + i1), &a0, &a1);
    jc_load_x2<PAIRS>(d, (size_t)(r10 + i), (size_t)(r10 + i1), &b0, &b1);
    jc_load_x2<PAIRS>(d, (size_t)(r01 + i), (size_t)(r01 + i1), &c0, &c1);
    jc_load_x2<PAIRS>(d, (size_t)(r11 + i), (size_t)(r11 + i1), &d0, &d1);
    const double a = jc_lerp(jc_lerp(a0, a1, fx), jc_lerp(b0, b1, fx), fy);
    const double b = jc_lerp(jc_lerp(c0, c1, fx), jc_lerp(d0, d1, fx), fy);
    return jc_lerp(a, b, fz);
}

// Corner cache of the advection kernel.  The 10 samples of one particle's midpoint sub-steps (pic_solver3.cpp:206-282) lie
// within a fraction of a cell of each other, so consecutive samples of a component mostly interpolate between the SAME 8
// corner values.  ncu (profiles/r02_c7_move.md) shows the kernel bound by L1 data-stage wavefronts — 120 128-bit loads per
// particle, 86 % of the LSU data pipe, FP64 pipe 40 % — so the corners of the last sample stay in registers, keyed by the
// linear index of the base corner, and are reloaded only when the base changes.  Same doubles, same expressions:
// bit-identical to the uncached form (tests/hostcheck runs this very source against the oracle).
template <typename I>
struct CornerCache {
    I key[3];
    double c[3][8];
    __host__ __device__ __forceinline__ CornerCache() { key[0] = key[1] = key[2] = (I)-1; }
};

template <typename I>
__host__ __device__ __forceinline__ double sample_rows_cached(const double* __restrict__ d, I r00, I r10, I r01, I r11, int i,
                                                              double fx, double fy, double fz, I* key, double* c) {
    const I k0 = r00 + (I)i;
    if (k0 != *key) {
        jc_load_x2<true>(d, (size_t)(r00 + i), 0, &c[0], &c[1]);
        jc_load_x2<true>(d, (size_t)(r10 + i), 0, &c[2], &c[3]);
        jc_load_x2<true>(d, (size_t)(r01 + i), 0, &c[4], &c[5]);
        jc_load_x2<true>(d, (size_t)(r11 + i), 0, &c[6], &c[7]);
        *key = k0;
    }
    const double a = jc_lerp(jc_lerp(c[0], c[1], fx), jc_lerp(c[2], c[3], fx), fy);
    const double b = jc_lerp(jc_lerp(c[4], c[5], fx), jc_lerp(c[6], c[7], fx), fy);
    return jc_lerp(a, b, fz);
}

template <bool POW2, typename I, bool PAIRS = false>
__host__ __device__ __forceinline__ void sample_velocity(const GridD& g, const double* __restrict__ u, const double* __restrict__ v,
                                                const double* __restrict__ w, double px, double py, double pz, double* ou,
                                                double* ov, double* ow, CornerCache<I>* cc = nullptr) {
    // FaceCenteredGrid3::sample: three LinearArraySampler3::operator() calls (face_centered_grid3.cpp:447-467).
    // The three component grids share their per-axis normalisations pairwise (node-aligned along the component's own
    // axis, cell-centred along the other two), so 6 barycentric evaluations serve all 3 samplers instead of 9; each is
    // the same expression the per-grid sampler evaluates, so the results are bit-identical to three jc_sample calls.
    int inx, icx, jny, jcy, knz, kcz;
    double fnx, fcx, fny, fcy, fnz, fcz;
    jc_barycentric(norm_pos<POW2>(px, g.ox + 0.5 * 0.0, g.hx, g.ihx), g.nx, &inx, &fnx);            // u grid: nx + 1 nodes
    jc_barycentric(norm_pos<POW2>(px, g.ox + 0.5 * g.hx, g.hx, g.ihx), g.nx - 1, &icx, &fcx);      // v, w grids
    jc_barycentric(norm_pos<POW2>(py, g.oy + 0.5 * 0.0, g.hy, g.ihy), g.ny, &jny, &fny);            // v grid
    jc_barycentric(norm_pos<POW2>(py, g.oy + 0.5 * g.hy, g.hy, g.ihy), g.ny - 1, &jcy, &fcy);      // u, w grids
    jc_barycentric(norm_pos<POW2>(pz, g.oz + 0.5 * 0.0, g.hz, g.ihz), g.nz, &knz, &fnz);            // w grid
    jc_barycentric(norm_pos<POW2>(pz, g.oz + 0.5 * g.hz, g.hz, g.ihz), g.nz - 1, &kcz, &fcz);      // u, v grids
    const int inx1 = min(inx + 1, g.nx), icx1 = min(icx + 1, g.nx - 1);
    const int jny1 = min(jny + 1, g.ny), jcy1 = min(jcy + 1, g.ny - 1);
    const int knz1 = min(knz + 1, g.nz), kcz1 = min(kcz + 1, g.nz - 1);
    if (PAIRS && cc) {   // the pair entry at i already holds the value at min(i + 1, sx - 1): i1 is not needed
        {
            const I sy = (I)(g.nx + 1), sz = sy * (I)g.ny;
            *ou = sample_rows_cached<I>(u, (I)kcz * sz + (I)jcy * sy, (I)kcz * sz + (I)jcy1 * sy, (I)kcz1 * sz + (I)jcy * sy,
                                        (I)kcz1 * sz + (I)jcy1 * sy, inx, fnx, fcy, fcz, &cc->key[0], cc->c[0]);
        }
        {
            const I sy = (I)g.nx, sz = sy * (I)(g.ny + 1);
            *ov = sample_rows_cached<I>(v, (I)kcz * sz + (I)jny * sy, (I)kcz * sz + (I)jny1 * sy, (I)kcz1 * sz + (I)jny * sy,
                                        (I)kcz1 * sz + (I)jny1 * sy, icx, fcx, fny, fcz, &cc->key[1], cc->c[1]);
        }
        {
            const I sy = (I)g.nx, sz = sy * (I)g.ny;
            *ow = sample_rows_cached<I>(w, (I)knz * sz + (I)jcy * sy, (I)knz * sz + (I)jcy1 * sy, (I)knz1 * sz + (I)jcy * sy,
                                        (I)knz1 * sz + (I)jcy1 * sy, icx, fcx, fcy, fnz, &cc->key[2], cc->c[2]);
        }
        return;
    }
    {
        const I sy = (I)(g.nx + 1), sz = sy * (I)g.ny;
        *ou = sample_rows<PAIRS, I>(u, (I)kcz * sz + (I)jcy * sy, (I)kcz * sz + (I)jcy1 * sy, (I)kcz1 * sz + (I)jcy * sy,
                                    (I)kcz1 * sz + (I)jcy1 * sy, inx, inx1, fnx, fcy, fcz);
    }
    {
        const I sy = (I)g.nx, sz = sy * (I)(g.ny + 1);
        *ov = sample_rows<PAIRS, I>(v, (I)kcz * sz + (I)jny * sy, (I)kcz * sz + (I)jny1 * sy, (I)kcz1 * sz + (I)jny * sy,
                                    (I)kcz1 * sz + (I)jny1 * sy, icx, icx1, fcx, fny, fcz);
    }
    {
        const I sy = (I)g.nx, sz = sy * (I)g.ny;
        *ow = sample_rows<PAIRS, I>(w, (I)knz * sz + (I)jcy * sy, (I)knz * sz + (I)jcy1 * sy, (I)knz1 * sz + (I)jcy * sy,
                                    (I)knz1 * sz + (I)jcy1 * sy, icx, icx1, fcx, fcy, fnz);
    }
}

// One entry of a pair array (see jc_load_x2): i = linear index into an array whose x extent is sx.
__host__ __device__ __forceinline__ void pack_pair(const double* __restrict__ d, double* __restrict__ out, size_t i, int sx) {
    const double a = d[i];
    const double b = ((int)(i % (size_t)sx) + 1 < sx) ? d[i + 1] : a;
    out[2 * i] = a; out[2 * i + 1] = b;
}


// FlipSolver3::transferFromGridsToParticles, src/jet/flip_solver3.cpp:50-99: v += sampleF32(delta)(float(x)),
// optional lerp towards the PIC sample.  pos: particle position, vel: in/out.
__host__ __device__ __forceinline__ void g2p_flip_particle(const GridD& g, const double* u, const double* v, const double* w,
                                                           const float* ud, const float* vd, const float* wd, double blend,
                                                           const double* pos, double* vel) {
    const double px = pos[0], py = pos[1], pz = pos[2];
    float xf = (float)px, yf = (float)py, zf = (float)pz;
    float fhx = (float)g.hx, fhy = (float)g.hy, fhz = (float)g.hz;
    GridView vu = comp_view(g, 0, u), vv = comp_view(g, 1, v), vw = comp_view(g, 2, w);
    // float(h) of a power-of-two h is the same power of two, so the exact-reciprocal shortcut carries over
    const float fix = (float)g.ihx, fiy = (float)g.ihy, fiz = (float)g.ihz;
    double du = jc_samplef(ud, vu.nx, vu.ny, vu.nz, fhx, fhy, fhz, (float)vu.ox, (float)vu.oy, (float)vu.oz, xf, yf, zf, fix, fiy, fiz);
    double dv = jc_samplef(vd, vv.nx, vv.ny, vv.nz, fhx, fhy, fhz, (float)vv.ox, (float)vv.oy, (float)vv.oz, xf, yf, zf, fix, fiy, fiz);
    double dw = jc_samplef(wd, vw.nx, vw.ny, vw.nz, fhx, fhy, fhz, (float)vw.ox, (float)vw.oy, (float)vw.oz, xf, yf, zf, fix, fiy, fiz);
    double f0 = vel[0] + du, f1 = vel[1] + dv, f2 = vel[2] + dw;
    if (blend > 0.0) {
        double pu = jc_sample(vu, px, py, pz), pv = jc_sample(vv, px, py, pz), pw = jc_sample(vw, px, py, pz);
        f0 = (1 - blend) * f0 + blend * pu;
        f1 = (1 - blend) * f1 + blend * pv;
        f2 = (1 - blend) * f2 + blend * pw;
    }
    vel[0] = f0; vel[1] = f1; vel[2] = f2;
}

// ApicSolver3::transferFromGridsToParticles, src/jet/apic_solver3.cpp:148-229: PIC sample + c = sum gradW * value
// with getCoordinatesAndGradientWeights (array_samplers3-inl.h:195-261) at the tangentially clamped position.
template <int AXIS>
__host__ __device__ __forceinline__ void apic_affine(const GridD& g, const double* __restrict__ data, double px, double py,
                                            double pz, double* c0, double* c1, double* c2) {
    int sx, sy, sz;
    double ox, oy, oz;
    comp_grid<AXIS>(g, &sx, &sy, &sz, &ox, &oy, &oz);
    const double hix = g.ox + g.hx * (double)g.nx, hiy = g.oy + g.hy * (double)g.ny, hiz = g.oz + g.hz * (double)g.nz;
    if (AXIS != 0) px = jc_clamp(px, g.ox + g.hx / 2.0, hix - g.hx / 2.0);
    if (AXIS != 1) py = jc_clamp(py, g.oy + g.hy / 2.0, hiy - g.hy / 2.0);
    if (AXIS != 2) pz = jc_clamp(pz, g.oz + g.hz / 2.0, hiz - g.hz / 2.0);
    BaseFrac b = jc_base_div(sx, sy, sz, g.hx, g.hy, g.hz, ox, oy, oz, px, py, pz, g.ihx, g.ihy, g.ihz);
    const double ix = 1.0 / g.hx, iy = 1.0 / g.hy, iz = 1.0 / g.hz;
    const double fx = b.fx, fy = b.fy, fz = b.fz, ax = 1 - b.fx, ay = 1 - b.fy, az = 1 - b.fz;
    double gw[8][3];
    gw[0][0] = -ix * ay * az;    gw[0][1] = -iy * ax * az;    gw[0][2] = -iz * ax * ay;
    gw[1][0] = ix * ay * az;     gw[1][1] = fx * (-iy) * az;  gw[1][2] = fx * ay * (-iz);
    gw[2][0] = (-ix) * fy * az;  gw[2][1] = ax * iy * az;     gw[2][2] = ax * fy * (-iz);
    gw[3][0] = ix * fy * az;     gw[3][1] = fx * iy * az;     gw[3][2] = fx * fy * (-iz);
    gw[4][0] = (-ix) * ay * fz;  gw[4][1] = ax * (-iy) * fz;  gw[4][2] = ax * ay * iz;
    gw[5][0] = ix * ay * fz;     gw[5][1] = fx * (-iy) * fz;  gw[5][2] = fx * ay * iz;
    gw[6][0] = (-ix) * fy * fz;  gw[6][1] = ax * iy * fz;     gw[6][2] = ax * fy * iz;
    gw[7][0] = ix * fy * fz;     gw[7][1] = fx * iy * fz;     gw[7][2] = fx * fy * iz;
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        int ii = (q & 1) ? b.ip1 : b.i, jj = (q & 2) ? b.jp1 : b.j, kk = (q & 4) ? b.kp1 : b.k;
        double val = data[((size_t)kk * sy + jj) * sx + ii];
        r0 += gw[q][0] * val; r1 += gw[q][1] * val; r2 += gw[q][2] * val;
    }
    *c0 = r0; *c1 = r1; *c2 = r2;
}


// PicSolver3::moveParticles, src/jet/pic_solver3.cpp:206-282: uint(max(maxCfl,1)) midpoint-rule sub-steps, then
// clamp to the domain box and zero the velocity component on closed walls (<= / >= tests).  pos, vel: in/out.
// first != nullptr: the grid velocity at the starting position is already known (a PIC transfer just sampled it there: the
// fused G2P + advection kernel) and replaces the first sample of the first sub-step — the same expression at the same point.
template <bool POW2, typename I, bool PAIRS = false>
__host__ __device__ __forceinline__ void move_particle(const GridD& g, const double* u, const double* v, const double* w, double dt,
                                                       unsigned nsub, int flag, double* pos, double* vel, const double* first = nullptr,
                                                       CornerCache<I>* cc = nullptr) {
    double p0x = pos[0], p0y = pos[1], p0z = pos[2];
    double p1x = p0x, p1y = p0y, p1z = p0z;
    for (unsigned t = 0; t < nsub; ++t) {
        double v0x, v0y, v0z, mvx, mvy, mvz;
        if (t == 0 && first) { v0x = first[0]; v0y = first[1]; v0z = first[2]; }
        else sample_velocity<POW2, I, PAIRS>(g, u, v, w, p0x, p0y, p0z, &v0x, &v0y, &v0z, cc);
        double mx = p0x + 0.5 * dt * v0x, my = p0y + 0.5 * dt * v0y, mz = p0z + 0.5 * dt * v0z;
        sample_velocity<POW2, I, PAIRS>(g, u, v, w, mx, my, mz, &mvx, &mvy, &mvz, cc);
        p1x = p0x + dt * mvx; p1y = p0y + dt * mvy; p1z = p0z + dt * mvz;
        p0x = p1x; p0y = p1y; p0z = p1z;
    }
    const double hix = g.ox + g.hx * (double)g.nx, hiy = g.oy + g.hy * (double)g.ny, hiz = g.oz + g.hz * (double)g.nz;
    double vx = vel[0], vy = vel[1], vz = vel[2];
    if ((flag & JETCU_DIR_LEFT) && p1x <= g.ox) { p1x = g.ox; vx = 0.0; }
    if ((flag & JETCU_DIR_RIGHT) && p1x >= hix) { p1x = hix; vx = 0.0; }
    if ((flag & JETCU_DIR_DOWN) && p1y <= g.oy) { p1y = g.oy; vy = 0.0; }
    if ((flag & JETCU_DIR_UP) && p1y >= hiy) { p1y = hiy; vy = 0.0; }
    if ((flag & JETCU_DIR_BACK) && p1z <= g.oz) { p1z = g.oz; vz = 0.0; }
    if ((flag & JETCU_DIR_FRONT) && p1z >= hiz) { p1z = hiz; vz = 0.0; }
    pos[0] = p1x; pos[1] = p1y; pos[2] = p1z;
    vel[0] = vx; vel[1] = vy; vel[2] = vz;
}
