// p2g_ops.cuh — arithmetic of the particle-to-grid transfer as __host__ __device__ functions: the per-cell register
// accumulation, the per-face gather of a tile's partial sums and the normalisation.  hybrid.cu's kernel is a wrapper
// (shared memory, thread loops, __syncthreads around them); tests/hostcheck runs the same source on the CPU against the
// oracle.
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


// ---- P2G: per-cell partial sums, exchanged through shared memory, gathered per face -----------------------------------
// transferFromParticlesToGrids: PIC src/jet/pic_solver3.cpp:118-193, APIC src/jet/apic_solver3.cpp:26-146.
//
// The reference accumulates `u += v*w; wsum += w` over the particles in one serial loop, so a parallel version has to choose
// its own summation order; ours is fixed by the GRID, never by scheduling (no atomics anywhere):
//   1. particles are cell-sorted (ascending original index inside a cell).  All particles of one cell deposit into the
//      same window of faces per component: 2 along the component's own (node-aligned) axis x 3 x 3 along the two
//      cell-centred axes = 18 faces.  One thread owns a cell and sums its particles' contributions to those 18 faces in
//      REGISTERS, in the cell's particle order (p2g_cell_partial);
//   2. the 18 (value, weight) partial sums of every cell of a tile go to shared memory, and one thread per FACE of the
//      tile's window adds up the partial sums of the <= 18 cells of the tile that reach it, in a fixed slot order
//      (p2g_face_gather);
//   3. faces on a tile's rim also receive from neighbouring tiles.  Tiles are launched in 8 colour passes (parity of the
//      tile index per axis); the windows of two tiles of one colour never overlap, so a tile adds its face sums into the
//      global arrays with a plain read-modify-write, and a rim face receives its <= 8 tile sums in colour order.
// Every face therefore sees one fixed sequence of FP64 additions that depends only on the particle set and on the
// globally aligned tile lattice — bit-reproducible run to run and across z-slab decompositions (a rank holds every cell
// within two planes of its slab, i.e. every cell that reaches one of its faces).  Versus the reference only the summation
// order differs (documented tolerance 1e-12 relative).
#define P2G_TX 8
#define P2G_TY 8
#define P2G_SLOTS 18
#ifndef P2G_TZ
#define P2G_TZ 8   // cells per tile along z (threads per CTA = P2G_TX * P2G_TY * P2G_TZ)
#endif
#ifndef P2G_TZ_APIC
#define P2G_TZ_APIC 4
#endif

// slot of window position (x, b, c) for component AX: the node axis has 2 positions, the centred axes 3
template <int AX>
__host__ __device__ constexpr int p2g_slot(int x, int b, int c) {
    return (c * (AX == 1 ? 2 : 3) + b) * (AX == 0 ? 2 : 3) + x;
}

// Where the cell loop reads a particle's attributes from.  Logical attribute ids: 0..2 position, 3 the velocity component
// of the pass, 4..6 that component's APIC affine vector.
#define P2G_A_VEL 3
#define P2G_A_C0 4
template <int AX>
struct P2GGlobalSrc {   // straight from the SoA arrays in HBM
    const AttrPtrs* a;
    __host__ __device__ __forceinline__ double get(int attr, uint32_t q) const {
        return a->p[attr < 3 ? attr : (attr == P2G_A_VEL ? 3 + AX : 6 + 3 * AX + (attr - P2G_A_C0))][q];
    }
};
// Partial sums of one cell for component AX.  sv / sw: P2G_SLOTS accumulators (value, weight); returns the mask of slots
// that are a real corner of at least one particle (the reference marks all 8 corners of a particle, also those that get
// weight 0).  Slots outside a particle's actual 2x2x2 corners receive weight 0 (val = +-0): adding them changes nothing.
template <bool APIC, int AX, class Src, bool PF = true>
__host__ __device__ __forceinline__ unsigned p2g_cell_partial(const GridD& g, int ci, int cj, int ck, uint32_t s0, uint32_t e0,
                                                              const Src& a, double* sv, double* sw, int* err) {
    constexpr int NX = AX == 0 ? 2 : 3, NY = AX == 1 ? 2 : 3, NZ = AX == 2 ? 2 : 3;
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    const double ox = g.ox + 0.5 * (AX == 0 ? 0.0 : g.hx), oy = g.oy + 0.5 * (AX == 1 ? 0.0 : g.hy),
                 oz = g.oz + 0.5 * (AX == 2 ? 0.0 : g.hz);
    const double ihx = 1.0 / g.hx, ihy = 1.0 / g.hy, ihz = 1.0 / g.hz;
    // window origin in face indices
    const int bx0 = ci - (AX == 0 ? 0 : 1), by0 = cj - (AX == 1 ? 0 : 1), bz0 = ck - (AX == 2 ? 0 : 1);
#pragma unroll
    for (int s = 0; s < P2G_SLOTS; ++s) { sv[s] = 0.0; sw[s] = 0.0; }
    unsigned touched = 0;
    // software pipeline: the next particle's attributes are requested before the current one is accumulated, so the
    // ~250-instruction dependent chain of a particle hides the L2 latency of the next loads (ncu, profiles/r02_p2g_variants.md:
    // long-scoreboard stalls were 43 % of the cycles between two issues with 16 resident warps per SM)
    double npx = 0.0, npy = 0.0, npz = 0.0, npv = 0.0, nc0 = 0.0, nc1 = 0.0, nc2 = 0.0;
    if (PF && s0 < e0) {
        npx = a.get(0, s0); npy = a.get(1, s0); npz = a.get(2, s0); npv = a.get(P2G_A_VEL, s0);
        if (APIC) { nc0 = a.get(P2G_A_C0, s0); nc1 = a.get(P2G_A_C0 + 1, s0); nc2 = a.get(P2G_A_C0 + 2, s0); }
    }
    for (uint32_t q = s0; q < e0; ++q) {
        double px = npx, py = npy, pz = npz;
        double pv = npv;
        double c0 = nc0, c1 = nc1, c2 = nc2;
        if (!PF) {   // plain form: load, then accumulate
            px = a.get(0, q); py = a.get(1, q); pz = a.get(2, q);
        } else if (q + 1 < e0) {
            npx = a.get(0, q + 1); npy = a.get(1, q + 1); npz = a.get(2, q + 1); npv = a.get(P2G_A_VEL, q + 1);
            if (APIC) { nc0 = a.get(P2G_A_C0, q + 1); nc1 = a.get(P2G_A_C0 + 1, q + 1); nc2 = a.get(P2G_A_C0 + 2, q + 1); }
        }
        if (APIC) {  // tangential clamp to [lo + h/2, hi - h/2] (apic_solver3.cpp:76-84,94-102,112-120)
            if (AX != 0) px = jc_clamp(px, g.ox + g.hx / 2.0, (g.ox + g.hx * (double)g.nx) - g.hx / 2.0);
            if (AX != 1) py = jc_clamp(py, g.oy + g.hy / 2.0, (g.oy + g.hy * (double)g.ny) - g.hy / 2.0);
            if (AX != 2) pz = jc_clamp(pz, g.oz + g.hz / 2.0, (g.oz + g.hz * (double)g.nz) - g.hz / 2.0);
        }
        const BaseFrac bf = jc_base_mul(sx, sy, sz, ihx, ihy, ihz, ox, oy, oz, px, py, pz);
        if (!PF) {
            pv = a.get(P2G_A_VEL, q);
            if (APIC) { c0 = a.get(P2G_A_C0, q); c1 = a.get(P2G_A_C0 + 1, q); c2 = a.get(P2G_A_C0 + 2, q); }
        }
        const int dx = bf.i - bx0, dy = bf.j - by0, dz = bf.k - bz0;  // 0 on the node axis, 0 or 1 on the centred ones
        const bool inwin = (unsigned)dx <= (AX == 0 ? 0u : 1u) && (unsigned)dy <= (AX == 1 ? 0u : 1u) &&
                           (unsigned)dz <= (AX == 2 ? 0u : 1u);
        if (!inwin) {  // cannot happen for a particle binned into this cell by cell_coord (same expressions); never silent
            *err = 1;
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
                    sv[p2g_slot<AX>(x, b, c)] += val; sw[p2g_slot<AX>(x, b, c)] += wgt;
                }
        // the particle's 2x2x2 corner block inside the window: the block at (0,0,0) shifted to (dx,dy,dz)
        constexpr unsigned block = 1u | (1u << 1) | (1u << NX) | (1u << (NX + 1)) | (1u << (NX * NY)) | (1u << (NX * NY + 1)) |
                                   (1u << (NX * NY + NX)) | (1u << (NX * NY + NX + 1));
        touched |= block << ((dz * NY + dy) * NX + dx);
    }
    return touched;
}

// One face of component AX's window of a tile of TX x TY x TZ cells (window = tile + one ring on the centred axes, + 1
// on the node axis): local face coordinates (lx, ly, lz).  ex = exchange buffer [P2G_SLOTS][ncell] holding one kind of
// partial sum (values or weights), tm = touched masks per cell.  Adds up the partial sums of the tile's cells that reach
// the face, in a fixed order; returns whether any cell touched the face.
template <int AX, int TZ>
__host__ __device__ __forceinline__ bool p2g_face_gather(const double* ex, const unsigned* tm, int lx, int ly, int lz, double* out) {
    constexpr int NX = AX == 0 ? 2 : 3, NY = AX == 1 ? 2 : 3, NZ = AX == 2 ? 2 : 3;
    constexpr int NCELL = P2G_TX * P2G_TY * TZ;
    double sum = 0.0;
    bool any = false;
#pragma unroll
    for (int c = 0; c < NZ; ++c) {
        const int cz = lz - c;
        if (cz < 0 || cz >= TZ) continue;
#pragma unroll
        for (int b = 0; b < NY; ++b) {
            const int cy = ly - b;
            if (cy < 0 || cy >= P2G_TY) continue;
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                const int cx = lx - x;
                if (cx < 0 || cx >= P2G_TX) continue;
                const int cell = (cz * P2G_TY + cy) * P2G_TX + cx;
                const int s = p2g_slot<AX>(x, b, c);
                if ((tm[cell] >> s) & 1u) {
                    any = true;
                    sum += ex[s * NCELL + cell];
                }
            }
        }
    }
    *out = sum;
    return any;
}

// ---- P2G on the STAGGERED lattice: 8 accumulators per thread instead of 18 ---------------------------------------------
// All particles whose sampler base index for component AX is (I, J, K) deposit into exactly the 8 faces (I..I+1, J..J+1,
// K..K+1) with weights that are products of (1 - f, f) per axis — no window, no zero-weight slots, no conditional weights.
// These particles are the "staggered cell" (I, J, K) of the component: on the component's node axis it coincides with the
// owning cell, on the two cell-centred axes it spans the upper half of cell S and the lower half of cell S + 1 (clamped
// bases at the walls add the outermost half cells).  The binning therefore sorts the particles by (cell, OCTANT) — the
// half of the cell per axis, p2g_octant — and a thread owning a staggered cell walks the octant ranges of the <= 4 cells
// that make it up, in a fixed order (cells: first centred axis fastest; octants ascending; original index inside an octant).
// Summation order per face: the 8 staggered cells around it in (z, y, x) order inside a tile (p2g_stag_gather), tiles in
// colour order — fixed by the lattice alone, hence bit-reproducible and independent of a z-slab decomposition.
__host__ __device__ __forceinline__ int jc_ffs32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __ffs((int)v);
#else
    return __builtin_ffs((int)v);
#endif
}

// Half of its cell a particle lies in, per axis, as the sampler of a cell-centred axis sees it: 1 iff the base index of
// that axis (floor of the position normalised against the data origin o + h/2, the expression of jc_base_mul) reaches the
// cell index.  Octant = (zh * 2 + yh) * 2 + xh.
__host__ __device__ __forceinline__ int p2g_octant(const GridD& g, double px, double py, double pz, int ci, int cj, int ck) {
    const int xh = floor((px - (g.ox + 0.5 * g.hx)) * (1.0 / g.hx)) >= (double)ci ? 1 : 0;
    const int yh = floor((py - (g.oy + 0.5 * g.hy)) * (1.0 / g.hy)) >= (double)cj ? 1 : 0;
    const int zh = floor((pz - (g.oz + 0.5 * g.hz)) * (1.0 / g.hz)) >= (double)ck ? 1 : 0;
    return (zh * 2 + yh) * 2 + xh;
}

// staggered cells along an axis of n cells: n on the component's node axis, n - 1 (at least 1) on a centred axis
__host__ __device__ __forceinline__ int p2g_stag_count(int n, bool node_axis) { return node_axis ? n : (n > 1 ? n - 1 : 1); }

// Accepted halves (bit 0: lower, bit 1: upper) of the cells S and S + 1 that make up staggered index S on a centred axis.
__host__ __device__ __forceinline__ void p2g_stag_halves(int S, int n, unsigned* hA, unsigned* hB) {
    *hA = 2u | (S == 0 ? 1u : 0u);                       // cell S: its upper half (+ the lower one at the low wall)
    *hB = S + 1 < n ? (1u | (S + 1 == n - 1 ? 2u : 0u)) : 0u;   // cell S + 1: its lower half (+ the upper one at the high wall)
    if (n == 1) *hA = 3u;
}

// Walks the particles of one staggered cell: the runs of accepted octants (bit 8 * c + o of `pend`) of its <= 4 cells, in
// ascending (cell, octant) order.  sub_start: start of every (cell, octant) range in the sorted particle arrays, 8 * cell + octant.
struct P2GWalker {
    const uint32_t* sub_start;
    size_t cell0, strideA, strideB;
    uint32_t pend, q, e;
    __host__ __device__ __forceinline__ uint32_t next() {   // index of the next particle; 0xffffffff when there is none
        while (q == e) {   // next non-empty run of accepted octants of one cell
            if (!pend) return 0xffffffffu;
            const int lo = jc_ffs32(pend) - 1;
            const uint32_t inv = ~(pend >> lo);
            int run = inv ? jc_ffs32(inv) - 1 : 8;
            if (run > 8 - (lo & 7)) run = 8 - (lo & 7);
            pend &= ~(((1u << run) - 1u) << lo);
            const int c = lo >> 3;
            const size_t sb = 8 * (cell0 + (size_t)(c & 1) * strideA + (size_t)(c >> 1) * strideB) + (size_t)(lo & 7);
            q = sub_start[sb]; e = sub_start[sb + run];
        }
        return q++;
    }
};

// Fraction of jc_barycentric(x, iHigh) (sampling.cuh) for a particle whose base index is known to be S, without branches
// and without the float-to-int conversion: with s = floor(x), the base is S iff s == S, or s < S at the low wall (S == 0:
// jc_barycentric clamps to i = 0, f = 0), or s > S at the high wall (S == iHigh - 1: i = iHigh - 1, f = 1).  *ok is
// cleared otherwise.  iHigh == 0 (a single face along the axis): i = 0, f = 0 whatever x is.
struct P2GAxis {
    double dS;
    bool low_ok, high_ok, deg;
    __host__ __device__ __forceinline__ void set(int S, int iHigh) {
        dS = (double)S; deg = iHigh == 0; low_ok = S == 0 || deg; high_ok = S == iHigh - 1 || deg;
    }
    __host__ __device__ __forceinline__ double frac(double x, bool* ok) const {
        const double s = floor(x);
        const bool lt = s < dS, gt = s > dS;
        double f = x - s;
        f = lt ? 0.0 : f; f = gt ? 1.0 : f; f = deg ? 0.0 : f;
        *ok &= !((lt && !low_ok) || (gt && !high_ok));
        return f;
    }
};

// Partial sums of staggered cell (I, J, K) of component AX: sv / sw = 8 accumulators (value, weight), slot (c * 2 + b) * 2 + x
// for face (I + x, J + b, K + c).  Cell planes outside [kb, ke) hold no particles on this rank.  Returns whether the
// staggered cell holds a particle (the reference marks all 8 corners of a particle, pic_solver3.cpp:160-176).  The next
// particle's attributes are requested before the current one is accumulated (the loads are one sector per lane: latency).
template <bool APIC, int AX, class Src>
__host__ __device__ __forceinline__ bool p2g_stag_partial(const GridD& g, int I, int J, int K, int kb, int ke,
                                                          const uint32_t* __restrict__ sub_start, const Src& a, double* sv, double* sw,
                                                          int* err) {
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    const double ox = g.ox + 0.5 * (AX == 0 ? 0.0 : g.hx), oy = g.oy + 0.5 * (AX == 1 ? 0.0 : g.hy),
                 oz = g.oz + 0.5 * (AX == 2 ? 0.0 : g.hz);
    const double ihx = 1.0 / g.hx, ihy = 1.0 / g.hy, ihz = 1.0 / g.hz;
#pragma unroll
    for (int s = 0; s < 8; ++s) { sv[s] = 0.0; sw[s] = 0.0; }
    // the <= 4 cells: A = first centred axis (cell-index stride strideA), B = second; octant masks per cell
    unsigned hx[2] = {3u, 0u}, hy[2] = {3u, 0u}, hz[2] = {3u, 0u};
    if (AX != 0) p2g_stag_halves(I, g.nx, &hx[0], &hx[1]);
    if (AX != 1) p2g_stag_halves(J, g.ny, &hy[0], &hy[1]);
    if (AX != 2) p2g_stag_halves(K, g.nz, &hz[0], &hz[1]);
    P2GWalker wk;
    wk.pend = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int da = c & 1, db = c >> 1;
        const int dx = AX == 0 ? 0 : da, dy = AX == 0 ? da : (AX == 1 ? 0 : db), dz = AX == 2 ? 0 : db;
        const int ck = K + dz;
        if (ck < kb || ck >= ke) continue;
        const unsigned mx = hx[dx], my = hy[dy], mz = hz[dz];
        const unsigned m8 = ((mx & 1u ? 0x55u : 0u) | (mx & 2u ? 0xAAu : 0u)) & ((my & 1u ? 0x33u : 0u) | (my & 2u ? 0xCCu : 0u)) &
                            ((mz & 1u ? 0x0Fu : 0u) | (mz & 2u ? 0xF0u : 0u));
        wk.pend |= m8 << (8 * c);
    }
    wk.sub_start = sub_start;
    wk.strideA = AX == 0 ? (size_t)g.nx : 1; wk.strideB = AX == 2 ? (size_t)g.nx : (size_t)g.nx * g.ny;
    wk.cell0 = ((size_t)K * g.ny + J) * g.nx + I;
    wk.q = 0; wk.e = 0;
    uint32_t cur = wk.next();
    if (cur == 0xffffffffu) return false;
    P2GAxis axx, axy, axz;
    axx.set(I, sx - 1); axy.set(J, sy - 1); axz.set(K, sz - 1);
    // face positions of the two corners per axis (APIC)
    const double gx0 = ox + g.hx * (double)I, gx1 = ox + g.hx * (double)(I + 1), gy0 = oy + g.hy * (double)J,
                 gy1 = oy + g.hy * (double)(J + 1), gz0 = oz + g.hz * (double)K, gz1 = oz + g.hz * (double)(K + 1);
    double npx = a.get(0, cur), npy = a.get(1, cur), npz = a.get(2, cur), npv = a.get(P2G_A_VEL, cur);
    double nc0 = 0.0, nc1 = 0.0, nc2 = 0.0;
    if (APIC) { nc0 = a.get(P2G_A_C0, cur); nc1 = a.get(P2G_A_C0 + 1, cur); nc2 = a.get(P2G_A_C0 + 2, cur); }
    bool ok = true;
    do {
        double px = npx, py = npy, pz = npz;
        const double pv = npv, c0 = nc0, c1 = nc1, c2 = nc2;
        cur = wk.next();
        if (cur != 0xffffffffu) {
            npx = a.get(0, cur); npy = a.get(1, cur); npz = a.get(2, cur); npv = a.get(P2G_A_VEL, cur);
            if (APIC) { nc0 = a.get(P2G_A_C0, cur); nc1 = a.get(P2G_A_C0 + 1, cur); nc2 = a.get(P2G_A_C0 + 2, cur); }
        }
        if (APIC) {  // tangential clamp to [lo + h/2, hi - h/2] (apic_solver3.cpp:76-84,94-102,112-120)
            if (AX != 0) px = jc_clamp(px, g.ox + g.hx / 2.0, (g.ox + g.hx * (double)g.nx) - g.hx / 2.0);
            if (AX != 1) py = jc_clamp(py, g.oy + g.hy / 2.0, (g.oy + g.hy * (double)g.ny) - g.hy / 2.0);
            if (AX != 2) pz = jc_clamp(pz, g.oz + g.hz / 2.0, (g.oz + g.hz * (double)g.nz) - g.hz / 2.0);
        }
        // base index + fractions of getCoordinatesAndWeights (array_samplers3-inl.h:151-193; jc_base_mul): the base is (I, J, K)
        // for a particle binned by cell_coord / p2g_octant — checked, never silent
        const double fx = axx.frac((px - ox) * ihx, &ok), fy = axy.frac((py - oy) * ihy, &ok), fz = axz.frac((pz - oz) * ihz, &ok);
        const double wx[2] = {1 - fx, fx}, wy[2] = {1 - fy, fy}, wz[2] = {1 - fz, fz};
        double ax[2], ay[2], az[2];
        if (APIC) {
            ax[0] = c0 * (gx0 - px); ax[1] = c0 * (gx1 - px);
            ay[0] = c1 * (gy0 - py); ay[1] = c1 * (gy1 - py);
            az[0] = c2 * (gz0 - pz); az[1] = c2 * (gz1 - pz);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    const double wgt = (wx[x] * wy[b]) * wz[c];  // (wx*wy)*wz as in array_samplers3-inl.h:176-192
                    const double val = APIC ? pv + ((ax[x] + ay[b]) + az[c]) : pv;
                    const int s = (c * 2 + b) * 2 + x;
                    sv[s] = fma(val, wgt, sv[s]); sw[s] += wgt;
                }
    } while (cur != 0xffffffffu);
    if (!ok) *err = 1;
    return true;
}

// One face (lx, ly, lz) of the window of a tile of TX x TY x TZ staggered cells (window = tile + 1 on every axis): adds up
// the partial sums of the tile's staggered cells around it in (z, y, x) order.  ex = [8][ncell] partial sums of one kind,
// tm = "holds a particle" per staggered cell.  Returns whether any of them does.
template <int TZ>
__host__ __device__ __forceinline__ bool p2g_stag_gather(const double* ex, const unsigned char* tm, int lx, int ly, int lz, double* out) {
    constexpr int NCELL = P2G_TX * P2G_TY * TZ;
    double sum = 0.0;
    bool any = false;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int cz = lz - c;
        if (cz < 0 || cz >= TZ) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int cy = ly - b;
            if (cy < 0 || cy >= P2G_TY) continue;
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const int cx = lx - x;
                if (cx < 0 || cx >= P2G_TX) continue;
                const int cell = (cz * P2G_TY + cy) * P2G_TX + cx;
                if (tm[cell]) {
                    any = true;
                    sum += ex[((c * 2 + b) * 2 + x) * NCELL + cell];
                }
            }
        }
    }
    *out = sum;
    return any;
}

// u = sum / wsum where wsum > 0 (pic_solver3.cpp:178-192); FLIP's float snapshot (src/jet/flip_solver3.cpp:39-47) is
// taken in the same pass.
__host__ __device__ __forceinline__ void p2g_normalise_face(double* u, const double* wsum, float* snapshot, size_t f) {
    const double w = wsum[f];
    double val = u[f];
    if (w > 0.0) val = val / w;
    u[f] = val;
    if (snapshot) snapshot[f] = (float)val;
}
