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

template <bool POW2, typename I>
__host__ __device__ __forceinline__ void sample_velocity(const GridD& g, const double* __restrict__ u, const double* __restrict__ v,
                                                const double* __restrict__ w, double px, double py, double pz, double* ou,
                                                double* ov, double* ow) {
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
    {
        const I sy = (I)(g.nx + 1), sz = sy * (I)g.ny;
        const I r00 = (I)kcz * sz + (I)jcy * sy, r10 = (I)kcz * sz + (I)jcy1 * sy, r01 = (I)kcz1 * sz + (I)jcy * sy,
                r11 = (I)kcz1 * sz + (I)jcy1 * sy;
        const double a = jc_lerp(jc_lerp(u[r00 + inx], u[r00 + inx1], fnx), jc_lerp(u[r10 + inx], u[r10 + inx1], fnx), fcy);
        const double b = jc_lerp(jc_lerp(u[r01 + inx], u[r01 + inx1], fnx), jc_lerp(u[r11 + inx], u[r11 + inx1], fnx), fcy);
        *ou = jc_lerp(a, b, fcz);
    }
    {
        const I sy = (I)g.nx, sz = sy * (I)(g.ny + 1);
        const I r00 = (I)kcz * sz + (I)jny * sy, r10 = (I)kcz * sz + (I)jny1 * sy, r01 = (I)kcz1 * sz + (I)jny * sy,
                r11 = (I)kcz1 * sz + (I)jny1 * sy;
        const double a = jc_lerp(jc_lerp(v[r00 + icx], v[r00 + icx1], fcx), jc_lerp(v[r10 + icx], v[r10 + icx1], fcx), fny);
        const double b = jc_lerp(jc_lerp(v[r01 + icx], v[r01 + icx1], fcx), jc_lerp(v[r11 + icx], v[r11 + icx1], fcx), fny);
        *ov = jc_lerp(a, b, fcz);
    }
    {
        const I sy = (I)g.nx, sz = sy * (I)g.ny;
        const I r00 = (I)knz * sz + (I)jcy * sy, r10 = (I)knz * sz + (I)jcy1 * sy, r01 = (I)knz1 * sz + (I)jcy * sy,
                r11 = (I)knz1 * sz + (I)jcy1 * sy;
        const double a = jc_lerp(jc_lerp(w[r00 + icx], w[r00 + icx1], fcx), jc_lerp(w[r10 + icx], w[r10 + icx1], fcx), fcy);
        const double b = jc_lerp(jc_lerp(w[r01 + icx], w[r01 + icx1], fcx), jc_lerp(w[r11 + icx], w[r11 + icx1], fcx), fcy);
        *ow = jc_lerp(a, b, fnz);
    }
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
template <bool POW2, typename I>
__host__ __device__ __forceinline__ void move_particle(const GridD& g, const double* u, const double* v, const double* w, double dt,
                                                       unsigned nsub, int flag, double* pos, double* vel) {
    double p0x = pos[0], p0y = pos[1], p0z = pos[2];
    double p1x = p0x, p1y = p0y, p1z = p0z;
    for (unsigned t = 0; t < nsub; ++t) {
        double v0x, v0y, v0z, mvx, mvy, mvz;
        sample_velocity<POW2, I>(g, u, v, w, p0x, p0y, p0z, &v0x, &v0y, &v0z);
        double mx = p0x + 0.5 * dt * v0x, my = p0y + 0.5 * dt * v0y, mz = p0z + 0.5 * dt * v0z;
        sample_velocity<POW2, I>(g, u, v, w, mx, my, mz, &mvx, &mvy, &mvz);
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
