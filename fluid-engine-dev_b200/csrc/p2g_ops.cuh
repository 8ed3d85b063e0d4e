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

// u = sum / wsum where wsum > 0 (pic_solver3.cpp:178-192); FLIP's float snapshot (src/jet/flip_solver3.cpp:39-47) is
// taken in the same pass.
__host__ __device__ __forceinline__ void p2g_normalise_face(double* u, const double* wsum, float* snapshot, size_t f) {
    const double w = wsum[f];
    double val = u[f];
    if (w > 0.0) val = val / w;
    u[f] = val;
    if (snapshot) snapshot[f] = (float)val;
}
