// hybrid.cu — device-resident PIC / FLIP / APIC solver (seam 3 of include/jetcu.h).
//
// Mirrors GridFluidSolver3::onAdvanceTimeStep (src/jet/grid_fluid_solver3.cpp:179-210) with the PicSolver3 /
// FlipSolver3 / ApicSolver3 stage overrides (src/jet/{pic,flip,apic}_solver3.cpp).  B200-first design:
//   * particles live in HBM as SoA doubles, CELL-SORTED every step by a one-pass counting sort (histogram ->
//     scan -> scatter -> per-cell order fix-up by original index), so that
//   * P2G is a deterministic GATHER per face over the 18 neighbouring cells' contiguous particle ranges
//     (no atomics on FP64 data; fixed summation order: cells z,y,x then ascending original index),
//   * the particle SDF is a pruned per-cell gather over the same ranges (exact min => bit-exact),
//   * G2P / advection read the grids with cell-coherent (L2-resident) access.
// Parity notes: everything except the P2G summation order (reference: global particle index order) replays
// the reference's expressions; P2G agrees to rounding (documented tolerance 1e-12 relative in the tests).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "particles.cuh"
#include "pressure.cuh"
#include "sampling.cuh"
#include "particle_ops.cuh"
#include "p2g_ops.cuh"
#include "diffusion.cuh"


struct jetcu_hybrid {
    jetcu_ctx* ctx = nullptr;
    jetcu_hybrid_params P{};
    int nx = 0, ny = 0, nz = 0;
    size_t nu = 0, nv = 0, nw = 0, nc = 0;
    double *u = nullptr, *v = nullptr, *w = nullptr;
    char *um = nullptr, *vm = nullptr, *wm = nullptr;
    char *valid0 = nullptr, *valid1 = nullptr;
    double* sdf = nullptr;
    float *ud = nullptr, *vd = nullptr, *wd = nullptr;
    double* pair[3] = {};                  // pair arrays of u, v, w for the advection samples (particle_ops.cuh: jc_load_x2)
    double* wsum[3] = {};                  // P2G weight sums (u, v, w faces); the value sums live in u, v, w themselves
    int* p2g_err = nullptr;                // device flag: a particle outside its cell's face window (cannot happen)
    // particles
    uint64_t np = 0, cap = 0;
    int nattr = 6;
    double* attr[2][NATTR] = {};
    uint32_t* orig[2] = {};
    int cur = 0;
    uint32_t *cell_of = nullptr, *rank = nullptr, *perm = nullptr;
    uint32_t* cell_start = nullptr;  // nc + 2: first particle of every cell (compact copy of sub_start[8 * cell])
    uint32_t* sub_start = nullptr;   // 8 * nc + 2: counts, then starts, of every (cell, octant) bin (p2g_ops.cuh: p2g_octant)
    bool binned = false;
    bool counted = false;  // cell_of / rank / per-cell counts already produced by the advection kernel of the last step
    // z-slab mode (the ctx had a multi-rank communicator when the solver was created): this rank owns the cell planes
    // [z0, z1); after binning attr[cur] holds [ghost-lo | owned | ghost-hi | dead], the owned block starts at `base`.
    int world = 1, me = 0, z0 = 0, z1 = 0;
    uint64_t base = 0;     // first owned particle inside attr[cur]
    uint64_t nsorted = 0;  // ghost-lo + owned + ghost-hi (the cell-sorted live range starts at 0)
    uint64_t nglobal = 0;  // particles on all ranks (== np on a single GPU)
    // analytic rigid collider; col.kind == 0: none
    ColliderD col;
    double* col_sdf = nullptr;                          // GridFractionalBoundaryConditionSolver3::_colliderSdf
    double *bnd_u = nullptr, *bnd_v = nullptr, *bnd_w = nullptr;  // collider velocity at the face positions (pressure rhs)
    double *tmp_u = nullptr, *tmp_v = nullptr, *tmp_w = nullptr;  // uTemp / vTemp / wTemp of constrainVelocity
    char *cm_u = nullptr, *cm_v = nullptr, *cm_w = nullptr;       // its uMarker / vMarker / wMarker
    PressureSystem ps;
    jetcu_mgpcg_result last{};
    double viscosity = 0.0;   // GridFluidSolver3::viscosityCoefficient(): > kEpsilonD switches computeViscosity on
    cudaEvent_t ev[12] = {};
};

static GridD gridd(const jetcu_hybrid* h) {
    const jetcu_grid& g = h->P.grid;
    return GridD{h->nx, h->ny, h->nz, g.hx, g.hy, g.hz, g.ox, g.oy, g.oz,
                 jc_pow2_inverse(g.hx), jc_pow2_inverse(g.hy), jc_pow2_inverse(g.hz)};
}

// ------------------------------------------------------------------------------------------------ cell sort
// Bin of a particle = 8 * cell + octant (the half of the cell per axis: the staggered-lattice P2G walks octant ranges).
// Particles whose cell plane lies outside [e0, e1) (slab mode: this rank's slab + ghost planes) go to the trash bin,
// 8 * nx*ny*nz, which sorts behind every real bin and is dropped.
__global__ void k_cell_count(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                             size_t np, GridD g, int e0, int e1, uint32_t* __restrict__ cell_of,
                             uint32_t* __restrict__ rank, uint32_t* __restrict__ count) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const double x = px[p], y = py[p], z = pz[p];
    int ci = cell_coord(x, g.ox, 1.0 / g.hx, g.nx);
    int cj = cell_coord(y, g.oy, 1.0 / g.hy, g.ny);
    int ck = cell_coord(z, g.oz, 1.0 / g.hz, g.nz);
    uint32_t bin = (uint32_t)((((size_t)ck * g.ny + cj) * g.nx + ci) * 8 + p2g_octant(g, x, y, z, ci, cj, ck));
    if (ck < e0 || ck >= e1) bin = (uint32_t)((size_t)g.nx * g.ny * g.nz * 8);
    cell_of[p] = bin;
    rank[p] = atomicAdd(&count[bin], 1u);  // slot inside the bin; order fixed up by k_sort_cells
}

__global__ void k_scatter_perm(const uint32_t* __restrict__ cell_of, const uint32_t* __restrict__ rank,
                               const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ perm, size_t np) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    perm[cell_start[cell_of[p]] + rank[p]] = (uint32_t)p;
}

// Make the order inside every bin deterministic: ascending ORIGINAL particle index (insertion sort; a bin holds about one
// particle, a cell about eight).  The thread of a cell's octant 0 also writes the compact per-cell start.
__global__ void k_sort_cells(const uint32_t* __restrict__ sub_start, const uint32_t* __restrict__ orig,
                             uint32_t* __restrict__ perm, uint32_t* __restrict__ cell_start, size_t nbins) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nbins) return;
    uint32_t s = sub_start[c];
    if ((c & 7) == 0) cell_start[c >> 3] = s;   // c == nbins: the end of the window
    if (c == nbins) return;
    uint32_t e = sub_start[c + 1];
    for (uint32_t a = s + 1; a < e; ++a) {
        uint32_t pa = perm[a];
        uint32_t ka = orig[pa];
        uint32_t b = a;
        while (b > s) {
            uint32_t pb = perm[b - 1];
            if (orig[pb] <= ka) break;
            perm[b] = pb;
            --b;
        }
        perm[b] = pa;
    }
}

// NA is a template parameter so that the attribute loop unrolls into compile-time indices of the kernel-parameter
// arrays: a run-time `nattr` makes the compiler copy both AttrPtrs to local memory (240 B of stack per thread — at
// 126.6 M particles ncu showed 31 GB of DRAM writes for 6.6 GB of payload, profiles/r02_c1_particle_kernels.md).
template <int NA>
__global__ void __launch_bounds__(256) k_gather_attrs(AttrPtrs src, AttrPtrs dst, const uint32_t* __restrict__ orig_src,
                                                      uint32_t* __restrict__ orig_dst, const uint32_t* __restrict__ perm, size_t np) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= np) return;
    const uint32_t from = perm[s];
    double v[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) v[a] = src.p[a][from];
    const uint32_t o = orig_src[from];
#pragma unroll
    for (int a = 0; a < NA; ++a) dst.p[a][s] = v[a];
    orig_dst[s] = o;
}

// ------------------------------------------------------------------------------------------------ z-slab routing
// Slab mode: rank q owns the cell planes [q*T, (q+1)*T) and additionally needs read-only GHOST copies
// of the particles within SLAB_GHOST planes of its slab (P2G reaches 1 plane, the particle SDF 2r = 1.7 h -> 2).
// After moveParticles every rank holds owned particles only; each one is copied to every OTHER rank whose extended
// slab contains its new plane (its new owner when it migrated, the neighbour when it sits near a slab face), and kept
// locally iff it lies in this rank's extended slab.  The cell sort that follows puts [ghost-lo | owned | ghost-hi] in
// order, so "owned" stays one contiguous range.  FILL = false only counts, FILL = true writes the index lists.
#define SLAB_GHOST JETCU_SLAB_GHOST_PLANES
// One atomicAdd per warp and destination (ballot + popc): a rank sends ~1 M particles to each neighbour per step and the
// per-particle atomics on the two or three counters serialised in the L2 — 0.95 ms for the counting pass and as much again
// for the fill at 2 GPUs, a cost that does not shrink with the rank count (JETCU_BIN_TRACE, profiles/r02_multi_b.md).
template <bool FILL>
__global__ void k_slab_route(const double* __restrict__ pz, size_t np, double oz, double ihz, int nz, int T, int world, int me,
                             uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ lists) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long m = 0;
    if (p < np) m = jc_slab_dest_mask(cell_coord(pz[p], oz, ihz, nz), T, world, me, SLAB_GHOST);
    const unsigned lane = threadIdx.x & 31u, full = 0xffffffffu;
    for (;;) {   // blockDim is a multiple of 32: whole warps take every iteration together
        const unsigned pending = __ballot_sync(full, m != 0ull);
        if (!pending) break;
        const int q = __shfl_sync(full, __ffsll((long long)m) - 1, __ffs((int)pending) - 1);   // a destination still pending
        const bool has = (m >> q) & 1ull;
        const unsigned b = __ballot_sync(full, has);
        uint32_t base = 0;
        const int leader = __ffs((int)b) - 1;
        if ((int)lane == leader) base = atomicAdd(&counts[q], (uint32_t)__popc(b));
        base = __shfl_sync(full, base, leader);
        if (has) {
            if (FILL) lists[offsets[q] + base + (uint32_t)__popc(b & ((1u << lane) - 1u))] = (uint32_t)p;
            m &= ~(1ull << q);
        }
    }
}

// message to one rank: (nattr + 1) blocks of `cnt` doubles (attributes, then the original index as a double)
template <int NA>
__global__ void k_slab_pack(AttrPtrs src, const uint32_t* __restrict__ orig, const uint32_t* __restrict__ list,
                            size_t cnt, double* __restrict__ msg) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const uint32_t p = list[i];
#pragma unroll
    for (int a = 0; a < NA; ++a) msg[(size_t)a * cnt + i] = src.p[a][p];
    msg[(size_t)NA * cnt + i] = (double)orig[p];
}

template <int NA>
__global__ void k_slab_unpack(const double* __restrict__ msg, size_t cnt, AttrPtrs dst, uint32_t* __restrict__ orig,
                              size_t at) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
#pragma unroll
    for (int a = 0; a < NA; ++a) dst.p[a][at + i] = msg[(size_t)a * cnt + i];
    orig[at + i] = (uint32_t)msg[(size_t)NA * cnt + i];
}

// ------------------------------------------------------------------------------------------------ P2G
// transferFromParticlesToGrids: PIC src/jet/pic_solver3.cpp:118-193, APIC src/jet/apic_solver3.cpp:26-146.
// Design and summation order: p2g_ops.cuh.  One CTA = one tile of P2G_TX x P2G_TY x TZ cells, one thread per cell;
// launched once per tile colour (parity of the tile index per axis), so a CTA owns every face its window reaches for
// the duration of the launch and adds into the global arrays without atomics.
// Measured alternatives (round 2, profiles/r02_p2g_variants.md): exchanging values and weights in two rounds through half
// the buffer (L1 hit rate 10 % -> 51 %, but two more barriers per component: 11.5 ms instead of 9.8 ms for the 8 launches)
// and staging the tile's particles in shared memory with coalesced row copies (14 ms) were both slower than this form.
template <bool APIC, int AX, int TZ, bool PF>
__device__ __forceinline__ void p2g_tile_component(const GridD& g, int I0, int J0, int K0, int ci, int cj, int ck, uint32_t s0,
                                                   uint32_t e0, const AttrPtrs& a, double* ex, unsigned* tm,
                                                   double* __restrict__ U, double* __restrict__ W, char* __restrict__ M, int* err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    constexpr int WX = P2G_TX + (AX == 0 ? 1 : 2), WY = P2G_TY + (AX == 1 ? 1 : 2), WZ = TZ + (AX == 2 ? 1 : 2);
    const int tid = threadIdx.x;
    unsigned touched = 0;
    if (e0 > s0) {
        double sv[P2G_SLOTS], sw[P2G_SLOTS];
        const P2GGlobalSrc<AX> src{&a};
        touched = p2g_cell_partial<APIC, AX, P2GGlobalSrc<AX>, PF>(g, ci, cj, ck, s0, e0, src, sv, sw, err);
#pragma unroll
        for (int s = 0; s < P2G_SLOTS; ++s) { ex[s * NT + tid] = sv[s]; ex[(P2G_SLOTS + s) * NT + tid] = sw[s]; }
    }
    tm[tid] = touched;
    __syncthreads();
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    const int fx0 = I0 - (AX == 0 ? 0 : 1), fy0 = J0 - (AX == 1 ? 0 : 1), fz0 = K0 - (AX == 2 ? 0 : 1);
    for (int f = tid; f < WX * WY * WZ; f += NT) {
        const int lx = f % WX, ly = (f / WX) % WY, lz = f / (WX * WY);
        const int fi = fx0 + lx, fj = fy0 + ly, fk = fz0 + lz;
        if ((unsigned)fi >= (unsigned)sx || (unsigned)fj >= (unsigned)sy || (unsigned)fk >= (unsigned)sz) continue;
        double v, w;
        if (p2g_face_gather<AX, TZ>(ex, tm, lx, ly, lz, &v)) {
            p2g_face_gather<AX, TZ>(ex + P2G_SLOTS * NT, tm, lx, ly, lz, &w);
            const size_t idx = ((size_t)fk * sy + fj) * sx + fi;
            U[idx] += v; W[idx] += w; M[idx] = 1;
        }
    }
    __syncthreads();
}

// prefetch.global.L2 of the window faces thread tid will add into (same face <-> thread mapping as the gather loop)
template <int AX, int TZ>
__device__ __forceinline__ void p2g_prefetch_faces(const GridD& g, int I0, int J0, int K0, const double* __restrict__ U,
                                                   const double* __restrict__ W) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    constexpr int WX = P2G_TX + (AX == 0 ? 1 : 2), WY = P2G_TY + (AX == 1 ? 1 : 2), WZ = TZ + (AX == 2 ? 1 : 2);
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    const int fx0 = I0 - (AX == 0 ? 0 : 1), fy0 = J0 - (AX == 1 ? 0 : 1), fz0 = K0 - (AX == 2 ? 0 : 1);
    for (int f = threadIdx.x; f < WX * WY * WZ; f += NT) {
        const int lx = f % WX, ly = (f / WX) % WY, lz = f / (WX * WY);
        const int fi = fx0 + lx, fj = fy0 + ly, fk = fz0 + lz;
        if ((unsigned)fi >= (unsigned)sx || (unsigned)fj >= (unsigned)sy || (unsigned)fk >= (unsigned)sz) continue;
        if (lx != 0 && (fi & 3) != 0) continue;   // one request per 32-byte sector of a face row is enough
        const size_t idx = ((size_t)fk * sy + fj) * sx + fi;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(U + idx));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(W + idx));
    }
}

struct P2GOut {
    double* val[3];   // u, v, w: value sums (normalised in place afterwards)
    double* wsum[3];  // weight sums
    char* marker[3];
};

#ifndef P2G_PF_DEFAULT
#define P2G_PF_DEFAULT 1
#endif
// MINB: CTAs per SM the register allocation is held to.  A 256-thread tile (TZ = 4) compiled without it takes 146 (APIC: 196)
// registers and ONE such CTA fits an SM — 8 warps; with MINB = 2 it is held to 128 and two tiles share an SM, so one tile's
// exchange / gather / read-modify-write phases overlap the other's particle loop.
// PF: software-pipelined particle loads in the cell loop (p2g_ops.cuh).
template <bool APIC, int TZ, int MINB, bool PF>
__global__ void __launch_bounds__(P2G_TX * P2G_TY * TZ, MINB)
    k_p2g_tile(GridD g, int par_x, int par_y, int tz_first, int kb, int ke, const uint32_t* __restrict__ cell_start, AttrPtrs a,
               P2GOut out, int* __restrict__ err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    extern __shared__ __align__(16) double p2g_ex[];  // [2][P2G_SLOTS][NT]: value partials, weight partials
    __shared__ unsigned tm[NT];
    const int tid = threadIdx.x;
    const int I0 = (2 * blockIdx.x + par_x) * P2G_TX, J0 = (2 * blockIdx.y + par_y) * P2G_TY, K0 = (tz_first + 2 * blockIdx.z) * TZ;
    const int ci = I0 + tid % P2G_TX, cj = J0 + (tid / P2G_TX) % P2G_TY, ck = K0 + tid / (P2G_TX * P2G_TY);
    uint32_t s0 = 0, e0 = 0;
    // cell planes [kb, ke): the whole grid, or (slab mode) this rank's slab plus its ghost planes
    if (ci < g.nx && cj < g.ny && ck >= kb && ck < ke) {
        const size_t cell = ((size_t)ck * g.ny + cj) * g.nx + ci;
        s0 = cell_start[cell]; e0 = cell_start[cell + 1];
    }
    if (!__syncthreads_or(e0 > s0)) return;
    if (PF) {   // the face sums are read-modify-written at the end of each component: start pulling them into L2 now
        p2g_prefetch_faces<0, TZ>(g, I0, J0, K0, out.val[0], out.wsum[0]);
        p2g_prefetch_faces<1, TZ>(g, I0, J0, K0, out.val[1], out.wsum[1]);
        p2g_prefetch_faces<2, TZ>(g, I0, J0, K0, out.val[2], out.wsum[2]);
    }
    p2g_tile_component<APIC, 0, TZ, PF>(g, I0, J0, K0, ci, cj, ck, s0, e0, a, p2g_ex, tm, out.val[0], out.wsum[0], out.marker[0], err);
    p2g_tile_component<APIC, 1, TZ, PF>(g, I0, J0, K0, ci, cj, ck, s0, e0, a, p2g_ex, tm, out.val[1], out.wsum[1], out.marker[1], err);
    p2g_tile_component<APIC, 2, TZ, PF>(g, I0, J0, K0, ci, cj, ck, s0, e0, a, p2g_ex, tm, out.val[2], out.wsum[2], out.marker[2], err);
}

// ---- P2G on the staggered lattice (default; the arithmetic and the summation order are described in p2g_ops.cuh) ------
// One CTA owns a tile of 8 x 8 x TZ STAGGERED cells of each component in turn, one thread per staggered cell: it walks the
// octant ranges that make up its staggered cell with 8 (value, weight) accumulator pairs in registers, the 16 partial sums go
// through shared memory once, one thread per face of the tile's 9 x 9 x (TZ + 1) window adds the <= 8 of them that reach
// it, and the tile's face sums are added into the global arrays by plain read-modify-write — tiles are launched in 8 colour
// passes whose windows never overlap.  Against the 18-slot window kernel above (k_p2g_tile, JETCU_P2G_KERNEL=window): no
// zero-weight slots (8 instead of 18 products per particle and component), 16 instead of 36 accumulators per thread (half
// the registers: twice the resident warps), an 8-term instead of an 18-term face gather through 32 KB instead of 73 KB of
// shared memory.
template <bool APIC, int AX, int TZ>
__device__ __forceinline__ void p2g_stag_component(const GridD& g, int I0, int J0, int K0, int kb, int ke,
                                                   const uint32_t* __restrict__ sub_start, const AttrPtrs& a, double* ex,
                                                   unsigned char* tm, double* __restrict__ U, double* __restrict__ W,
                                                   char* __restrict__ M, int* err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    constexpr int WX = P2G_TX + 1, WY = P2G_TY + 1, WZ = TZ + 1;
    const int tid = threadIdx.x;
    const int I = I0 + tid % P2G_TX, J = J0 + (tid / P2G_TX) % P2G_TY, K = K0 + tid / (P2G_TX * P2G_TY);
    const int NI = p2g_stag_count(g.nx, AX == 0), NJ = p2g_stag_count(g.ny, AX == 1), NK = p2g_stag_count(g.nz, AX == 2);
    bool any = false;
    if (I < NI && J < NJ && K < NK) {
        double sv[8], sw[8];
        const P2GGlobalSrc<AX> src{&a};
        any = p2g_stag_partial<APIC, AX, P2GGlobalSrc<AX>>(g, I, J, K, kb, ke, sub_start, src, sv, sw, err);
        if (any) {
#pragma unroll
            for (int s = 0; s < 8; ++s) { ex[s * NT + tid] = sv[s]; ex[(8 + s) * NT + tid] = sw[s]; }
        }
    }
    tm[tid] = any ? 1 : 0;
    if (__syncthreads_or(any)) {
        const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
        for (int f = tid; f < WX * WY * WZ; f += NT) {
            const int lx = f % WX, ly = (f / WX) % WY, lz = f / (WX * WY);
            const int fi = I0 + lx, fj = J0 + ly, fk = K0 + lz;
            if (fi >= sx || fj >= sy || fk >= sz) continue;
            double v, w;
            if (p2g_stag_gather<TZ>(ex, tm, lx, ly, lz, &v)) {
                p2g_stag_gather<TZ>(ex + 8 * NT, tm, lx, ly, lz, &w);
                const size_t idx = ((size_t)fk * sy + fj) * sx + fi;
                U[idx] += v; W[idx] += w; M[idx] = 1;
            }
        }
    }
    __syncthreads();
}

template <int AX, int TZ>
__device__ __forceinline__ void p2g_stag_prefetch(const GridD& g, int I0, int J0, int K0, const double* __restrict__ U,
                                                  const double* __restrict__ W) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    constexpr int WX = P2G_TX + 1, WY = P2G_TY + 1, WZ = TZ + 1;
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    for (int f = threadIdx.x; f < WX * WY * WZ; f += NT) {
        const int lx = f % WX, ly = (f / WX) % WY, lz = f / (WX * WY);
        const int fi = I0 + lx, fj = J0 + ly, fk = K0 + lz;
        if (fi >= sx || fj >= sy || fk >= sz) continue;
        if (lx != 0 && (fi & 3) != 0) continue;   // one request per 32-byte sector of a face row is enough
        const size_t idx = ((size_t)fk * sy + fj) * sx + fi;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(U + idx));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(W + idx));
    }
}

// MINB: CTAs per SM the register allocation is held to (3: 85 registers, 4: 64)
template <bool APIC, int TZ, int MINB>
__global__ void __launch_bounds__(P2G_TX * P2G_TY * TZ, MINB)
    k_p2g_stag(GridD g, int par_x, int par_y, int tz_first, int kb, int ke, const uint32_t* __restrict__ cell_start,
               const uint32_t* __restrict__ sub_start, AttrPtrs a, P2GOut out, int* __restrict__ err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    __shared__ double ex[16 * NT];   // [2][8][NT]: value partials, weight partials
    __shared__ unsigned char tm[NT];
    const int tid = threadIdx.x;
    const int I0 = (2 * blockIdx.x + par_x) * P2G_TX, J0 = (2 * blockIdx.y + par_y) * P2G_TY, K0 = (tz_first + 2 * blockIdx.z) * TZ;
    {   // a tile whose cells [I0, I0 + 8] x [J0, J0 + 8] x [K0, K0 + TZ] hold no particle deposits nothing: each thread looks
        // at the two x-neighbours of its cell in the 2 x 2 rows above it (compact per-cell starts, coalesced)
        const int ci = I0 + tid % P2G_TX, cj = J0 + (tid / P2G_TX) % P2G_TY, ck = K0 + tid / (P2G_TX * P2G_TY);
        bool have = false;
        if (ci < g.nx) {
            const int i1 = min(ci + 2, g.nx);
            for (int dz = 0; dz < 2; ++dz)
                for (int dy = 0; dy < 2; ++dy) {
                    const int j = cj + dy, k = ck + dz;
                    if (j >= g.ny || k < kb || k >= ke) continue;
                    const size_t row = ((size_t)k * g.ny + j) * g.nx;
                    have |= cell_start[row + i1] > cell_start[row + ci];
                }
        }
        if (!__syncthreads_or(have)) return;
    }
    p2g_stag_prefetch<0, TZ>(g, I0, J0, K0, out.val[0], out.wsum[0]);
    p2g_stag_prefetch<1, TZ>(g, I0, J0, K0, out.val[1], out.wsum[1]);
    p2g_stag_prefetch<2, TZ>(g, I0, J0, K0, out.val[2], out.wsum[2]);
    p2g_stag_component<APIC, 0, TZ>(g, I0, J0, K0, kb, ke, sub_start, a, ex, tm, out.val[0], out.wsum[0], out.marker[0], err);
    p2g_stag_component<APIC, 1, TZ>(g, I0, J0, K0, kb, ke, sub_start, a, ex, tm, out.val[1], out.wsum[1], out.marker[1], err);
    p2g_stag_component<APIC, 2, TZ>(g, I0, J0, K0, kb, ke, sub_start, a, ex, tm, out.val[2], out.wsum[2], out.marker[2], err);
}

// u = sum / wsum where wsum > 0 (pic_solver3.cpp:178-192); FLIP's float snapshot (src/jet/flip_solver3.cpp:39-47) is
// taken in the same pass.
__global__ void k_p2g_normalise(double* __restrict__ u, const double* __restrict__ wsum, float* __restrict__ snapshot, size_t n) {
    size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    p2g_normalise_face(u, wsum, snapshot, f);
}

__global__ void k_snapshot_f32(const double* __restrict__ in, float* __restrict__ d, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = (float)in[i];
}

__global__ void k_flip_delta(const double* __restrict__ in, float* __restrict__ d, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = (float)in[i] - d[i];  // src/jet/flip_solver3.cpp:57-70
}

// ------------------------------------------------------------------------------------------------ particle SDF
// PicSolver3::buildSignedDistanceField, src/jet/pic_solver3.cpp:296-317: phi(cell) = min(2r, min |x_c - x_p|) - r
// over particles with |x_p - x_c|^2 <= (2r)^2.  The reference finds them through a 64^3 hash grid of spacing 4r
// (8 buckets around the query); any superset of "all particles within 2r" yields the same exact min, so we walk
// the cell-sorted ranges within `reach` cells and prune by box distance.
__global__ void __launch_bounds__(128)
    k_build_sdf(GridD g, const uint32_t* __restrict__ cell_start, const double* __restrict__ ppx,
                const double* __restrict__ ppy, const double* __restrict__ ppz, double radius, double band, int rx, int ry,
                int rz, int kfirst, double* __restrict__ sdf) {
    int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = kfirst + blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    // data position = (origin + 0.5 h) + h*(i,j,k)  (src/jet/scalar_grid3.cpp:127-132)
    const double ptx = (g.ox + 0.5 * g.hx) + g.hx * (double)i;
    const double pty = (g.oy + 0.5 * g.hy) + g.hy * (double)j;
    const double ptz = (g.oz + 0.5 * g.hz) + g.hz * (double)k;
    const double r2 = band * band;
    double minDist = band;
    double best2 = r2;  // pruning bound on squared distance (only ever shrinks)
    // visit order: own cell first (tightens the bound fastest), then the rest; the min is order-independent
    for (int pass = 0; pass < 2; ++pass) {
        int k0 = pass == 0 ? k : max(k - rz, 0), k1 = pass == 0 ? k : min(k + rz, g.nz - 1);
        int j0 = pass == 0 ? j : max(j - ry, 0), j1 = pass == 0 ? j : min(j + ry, g.ny - 1);
        int i0 = pass == 0 ? i : max(i - rx, 0), i1 = pass == 0 ? i : min(i + rx, g.nx - 1);
        for (int ck = k0; ck <= k1; ++ck) {
            // conservative distance from pt to the slab of cells ck (open towards the domain exterior)
            double dz = 0.0;
            if (ck > k) dz = (g.oz + g.hz * (double)ck) - ptz; else if (ck < k) dz = ptz - (g.oz + g.hz * (double)(ck + 1));
            dz = fmax(dz, 0.0) * 0.999999;
            for (int cj = j0; cj <= j1; ++cj) {
                double dy = 0.0;
                if (cj > j) dy = (g.oy + g.hy * (double)cj) - pty; else if (cj < j) dy = pty - (g.oy + g.hy * (double)(cj + 1));
                dy = fmax(dy, 0.0) * 0.999999;
                if (dz * dz + dy * dy > best2) continue;
                size_t row = ((size_t)ck * g.ny + cj) * g.nx;
                for (int ci = i0; ci <= i1; ++ci) {
                    if (pass == 1 && ci == i && cj == j && ck == k) continue;
                    double dx = 0.0;
                    if (ci > i) dx = (g.ox + g.hx * (double)ci) - ptx; else if (ci < i) dx = ptx - (g.ox + g.hx * (double)(ci + 1));
                    dx = fmax(dx, 0.0) * 0.999999;
                    if (dz * dz + dy * dy + dx * dx > best2) continue;
                    uint32_t s = cell_start[row + ci], e = cell_start[row + ci + 1];
                    for (uint32_t q = s; q < e; ++q) {
                        double ex = ppx[q] - ptx, ey = ppy[q] - pty, ez = ppz[q] - ptz;
                        double d2 = ex * ex + ey * ey + ez * ez;
                        if (d2 <= r2) {
                            double fx = ptx - ppx[q], fy = pty - ppy[q], fz = ptz - ppz[q];
                            double dist = sqrt(fx * fx + fy * fy + fz * fz);
                            if (dist < minDist) { minDist = dist; best2 = fmin(best2, d2 * 1.000001); }
                        }
                    }
                }
            }
        }
    }
    sdf[((size_t)k * g.ny + j) * g.nx + i] = minDist - radius;
}

// ------------------------------------------------------------------------------------------------ extrapolation
// One sweep of extrapolateToRegion, include/jet/detail/array_utils-inl.h:191-239: reads only cells valid at sweep
// start, writes only invalid ones (so it is safe in place); neighbour order +i,-i,+j,-j,+k,-k; divide by count.
__global__ void __launch_bounds__(256) k_extrapolate_sweep(double* data, const char* __restrict__ v0, char* __restrict__ v1,
                                                           int nx, int ny, int nz) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= nx || j >= ny) return;
    extrapolate_cell(data, v0, v1, nx, ny, nz, i, j, k);
}

// ------------------------------------------------------------------------------------------------ grid ops
// Closed-wall part of GridFractionalBoundaryConditionSolver3::constrainVelocity
// (src/jet/grid_fractional_boundary_condition_solver3.cpp:190-232).  Without a collider the rest of
// constrainVelocity is the identity (all face markers 1, no cell inside the collider SDF = kMaxD).
__global__ void k_zero_walls(GridD g, int flag, double* __restrict__ u, double* __restrict__ v, double* __restrict__ w) {
    zero_walls_elem(g, flag, u, v, w, (size_t)blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void k_add_scalar(double* __restrict__ a, double s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += s;  // computeGravity: u += dt * g  (src/jet/grid_fluid_solver3.cpp:317-344)
}

// GridFluidSolver3::cfl, src/jet/grid_fluid_solver3.cpp:57-72: max over cells of the SIGNED components of
// valueAtCellCenter + dt*g, starting from 0 (max is exact, so any reduction order gives the same bits).
#define CFL_BLOCK 256
#define CFL_GRID 592
__global__ void __launch_bounds__(CFL_BLOCK) k_cfl_partial(GridD g, const double* __restrict__ u, const double* __restrict__ v,
                                                           const double* __restrict__ w, double dtgx, double dtgy, double dtgz,
                                                           double* __restrict__ partials) {
    __shared__ double sm[CFL_BLOCK];
    double m = 0.0;
    size_t nc = (size_t)g.nx * g.ny * g.nz;
    for (size_t c = (size_t)blockIdx.x * CFL_BLOCK + threadIdx.x; c < nc; c += (size_t)CFL_GRID * CFL_BLOCK) {
        int i = (int)(c % g.nx);
        size_t t = c / g.nx;
        int j = (int)(t % g.ny), k = (int)(t / g.ny);
        size_t iu = ((size_t)k * g.ny + j) * (g.nx + 1) + i;
        size_t iv = ((size_t)k * (g.ny + 1) + j) * g.nx + i;
        double vx = 0.5 * (u[iu] + u[iu + 1]) + dtgx;
        double vy = 0.5 * (v[iv] + v[iv + g.nx]) + dtgy;
        double vz = 0.5 * (w[c] + w[c + (size_t)g.nx * g.ny]) + dtgz;
        m = fmax(m, vx); m = fmax(m, vy); m = fmax(m, vz);
    }
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = CFL_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(CFL_BLOCK) k_max_final(const double* __restrict__ partials, int n, double* out) {
    __shared__ double sm[CFL_BLOCK];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += CFL_BLOCK) m = fmax(m, partials[i]);
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = CFL_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// ------------------------------------------------------------------------------------------------ G2P + advection
// PicSolver3::transferFromGridsToParticles, src/jet/pic_solver3.cpp:195-204.  (Per-particle bodies: particle_ops.cuh.)
template <bool POW2, typename I>
__global__ void k_g2p_pic(GridD g, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                          AttrPtrs a, size_t np) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double ou, ov, ow;
    sample_velocity<POW2, I>(g, u, v, w, a.p[0][p], a.p[1][p], a.p[2][p], &ou, &ov, &ow);
    a.p[3][p] = ou; a.p[4][p] = ov; a.p[5][p] = ow;
}

__global__ void k_g2p_flip(GridD g, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                           const float* __restrict__ ud, const float* __restrict__ vd, const float* __restrict__ wd,
                           double blend, AttrPtrs a, size_t np) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const double pos[3] = {a.p[0][p], a.p[1][p], a.p[2][p]};
    double vel[3] = {a.p[3][p], a.p[4][p], a.p[5][p]};
    g2p_flip_particle(g, u, v, w, ud, vd, wd, blend, pos, vel);
    a.p[3][p] = vel[0]; a.p[4][p] = vel[1]; a.p[5][p] = vel[2];
}

template <bool POW2, typename I>
__global__ void k_g2p_apic(GridD g, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                           AttrPtrs a, size_t np) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double px = a.p[0][p], py = a.p[1][p], pz = a.p[2][p];
    double ou, ov, ow;
    sample_velocity<POW2, I>(g, u, v, w, px, py, pz, &ou, &ov, &ow);
    a.p[3][p] = ou; a.p[4][p] = ov; a.p[5][p] = ow;
    double c0, c1, c2;
    apic_affine<0>(g, u, px, py, pz, &c0, &c1, &c2); a.p[6][p] = c0; a.p[7][p] = c1; a.p[8][p] = c2;
    apic_affine<1>(g, v, px, py, pz, &c0, &c1, &c2); a.p[9][p] = c0; a.p[10][p] = c1; a.p[11][p] = c2;
    apic_affine<2>(g, w, px, py, pz, &c0, &c1, &c2); a.p[12][p] = c0; a.p[13][p] = c1; a.p[14][p] = c2;
}

// Tail of the advection kernels on a single GPU without a collider: the particle's NEW cell, its slot inside the cell and
// the per-cell counts — k_cell_count's work for the next step's binning — while the final position is still in registers
// (the next step then skips k_cell_count and its 24 B/particle re-read of the positions).
__device__ __forceinline__ void count_cell(const GridD& g, const double* pos, size_t p, uint32_t* __restrict__ cell_of,
                                           uint32_t* __restrict__ rank, uint32_t* __restrict__ count) {
    if (!count) return;
    const int ci = cell_coord(pos[0], g.ox, 1.0 / g.hx, g.nx);
    const int cj = cell_coord(pos[1], g.oy, 1.0 / g.hy, g.ny);
    const int ck = cell_coord(pos[2], g.oz, 1.0 / g.hz, g.nz);
    const uint32_t bin = (uint32_t)((((size_t)ck * g.ny + cj) * g.nx + ci) * 8 + p2g_octant(g, pos[0], pos[1], pos[2], ci, cj, ck));
    cell_of[p] = bin;
    rank[p] = atomicAdd(&count[bin], 1u);
}

// CC: keep the 8 corner values per component of the last sample in registers (particle_ops.cuh: CornerCache)
// CC = 0: no cache, 5 CTAs of 128 threads per SM (96 registers); 1: cache, 4 CTAs (128 registers); 2: cache, 5 CTAs (96
// registers, 48 B of spills); 3 / 4: no cache, 6 / 8 CTAs (80 / 64 registers); 5: cache, 6 CTAs (80 registers, 172 B of spills).
// JETCU_MOVE_CACHE selects one for all three kernels; default: the fastest measured per kernel at 256^3 (GPU call 10:
// FLIP 18.5 / - / 17.7 / 17.8 / 17.8 / 16.9 ms for 0..5, PIC 16.6 / - / 14.7 / 16.1 / 15.9 / 15.7 ms).  Occupancy alone does
// nothing (variants 3, 4): the kernel is bound by its FP64 instruction count (profiles/r02_move_variants.md).
#define MOVE_MIN_BLOCKS(CC) ((CC) == 1 ? 4 : ((CC) == 3 || (CC) == 5) ? 6 : ((CC) == 4 ? 8 : 5))
#define MOVE_CACHED(CC) ((CC) == 1 || (CC) == 2 || (CC) == 5)
template <bool POW2, typename I, int CC>
__global__ void __launch_bounds__(128, MOVE_MIN_BLOCKS(CC))
    k_move(GridD g, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                       double dt, unsigned nsub, int flag, AttrPtrs a, size_t np, uint32_t* __restrict__ cell_of,
                       uint32_t* __restrict__ rank, uint32_t* __restrict__ count) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double pos[3] = {a.p[0][p], a.p[1][p], a.p[2][p]}, vel[3] = {a.p[3][p], a.p[4][p], a.p[5][p]};
    CornerCache<I> cc;
    move_particle<POW2, I, true>(g, u, v, w, dt, nsub, flag, pos, vel, nullptr, MOVE_CACHED(CC) ? &cc : nullptr);   // u, v, w: PAIR arrays (k_pack_pairs)
    a.p[0][p] = pos[0]; a.p[1][p] = pos[1]; a.p[2][p] = pos[2];
    count_cell(g, pos, p, cell_of, rank, count);
    a.p[3][p] = vel[0]; a.p[4][p] = vel[1]; a.p[5][p] = vel[2];
}

// Fused transferFromGridsToParticles + moveParticles of one sub-step (PicSolver3::computeAdvection runs them back to back,
// src/jet/pic_solver3.cpp:95-112): the particle arrays are read and written once instead of twice.
//   PIC : the grid velocity sampled at the particle is its new velocity AND the first sample of the first midpoint
//         sub-step (same expression, same point), so one of the 11 samples disappears as well;
//   FLIP: velocity += float delta sample (pic blending factor 0; a non-zero factor takes the two-kernel path).
// u, v, w are pair arrays.
template <bool POW2, typename I, int CC>
__global__ void __launch_bounds__(128, MOVE_MIN_BLOCKS(CC))
    k_pic_move(GridD g, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                       double dt, unsigned nsub, int flag, AttrPtrs a, size_t np, uint32_t* __restrict__ cell_of,
                       uint32_t* __restrict__ rank, uint32_t* __restrict__ count) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double pos[3] = {a.p[0][p], a.p[1][p], a.p[2][p]}, vel[3];
    CornerCache<I> cc;
    sample_velocity<POW2, I, true>(g, u, v, w, pos[0], pos[1], pos[2], &vel[0], &vel[1], &vel[2], MOVE_CACHED(CC) ? &cc : nullptr);
    const double first[3] = {vel[0], vel[1], vel[2]};
    move_particle<POW2, I, true>(g, u, v, w, dt, nsub, flag, pos, vel, first, MOVE_CACHED(CC) ? &cc : nullptr);
    a.p[0][p] = pos[0]; a.p[1][p] = pos[1]; a.p[2][p] = pos[2];
    count_cell(g, pos, p, cell_of, rank, count);
    a.p[3][p] = vel[0]; a.p[4][p] = vel[1]; a.p[5][p] = vel[2];
}

template <bool POW2, typename I, int CC>
__global__ void __launch_bounds__(128, MOVE_MIN_BLOCKS(CC))
    k_flip_move(GridD g, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                            const float* __restrict__ ud, const float* __restrict__ vd, const float* __restrict__ wd, double dt,
                            unsigned nsub, int flag, AttrPtrs a, size_t np, uint32_t* __restrict__ cell_of,
                            uint32_t* __restrict__ rank, uint32_t* __restrict__ count) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double pos[3] = {a.p[0][p], a.p[1][p], a.p[2][p]}, vel[3] = {a.p[3][p], a.p[4][p], a.p[5][p]};
    g2p_flip_particle(g, nullptr, nullptr, nullptr, ud, vd, wd, 0.0, pos, vel);
    CornerCache<I> cc;
    move_particle<POW2, I, true>(g, u, v, w, dt, nsub, flag, pos, vel, nullptr, MOVE_CACHED(CC) ? &cc : nullptr);
    a.p[0][p] = pos[0]; a.p[1][p] = pos[1]; a.p[2][p] = pos[2];
    count_cell(g, pos, p, cell_of, rank, count);
    a.p[3][p] = vel[0]; a.p[4][p] = vel[1]; a.p[5][p] = vel[2];
}

// pair array of one component (particle_ops.cuh: jc_load_x2); with d != nullptr also FLIP's delta = float(u) - snapshot
// (src/jet/flip_solver3.cpp:57-70) in the same pass over the field
__global__ void k_pack_pairs(const double* __restrict__ in, double* __restrict__ pairs, float* __restrict__ d, int sx, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pack_pair(in, pairs, i, sx);
    if (d) d[i] = (float)in[i] - d[i];
}

// ------------------------------------------------------------------------------------------------ particle I/O
__global__ void k_aos_to_soa3(const double* __restrict__ aos, double* __restrict__ x, double* __restrict__ y,
                              double* __restrict__ z, size_t n, size_t offset) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[offset + i] = aos[3 * i]; y[offset + i] = aos[3 * i + 1]; z[offset + i] = aos[3 * i + 2];
}

// sorted SoA -> caller's AoS in ORIGINAL particle order
__global__ void k_soa_to_aos3_orig(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                   const uint32_t* __restrict__ orig, double* __restrict__ aos, size_t n) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    size_t o = orig[s];
    aos[3 * o] = x[s]; aos[3 * o + 1] = y[s]; aos[3 * o + 2] = z[s];
}

__global__ void k_iota_u32(uint32_t* p, uint32_t start, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = start + (uint32_t)i;
}

__global__ void k_fill_f64(double* p, double v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// GridPointGenerator3::forEachPoint lattice, src/jet/grid_point_generator3.cpp:12-37: position = idx*spacing + lower,
// emitted k-outer, j, i-inner.  A point is kept iff SurfaceToImplicit3(Box3).signedDistance(p) <= 0
// (src/jet/volume_particle_emitter3.cpp:93), replayed literally because for points exactly ON a box face the outcome
// depends on the rounding of Plane3::closestPointLocal (src/jet/plane3.cpp:35-38): Box3::closestPointLocal /
// closestNormalLocal (src/jet/box3.cpp:27-99), Surface3::isInsideLocal (src/jet/surface3.cpp:75-79),
// SurfaceToImplicit3::signedDistanceLocal (src/jet/surface_to_implicit3.cpp:67-72).
struct Box3d {
    double lo[3], hi[3];
};

__device__ __forceinline__ bool box_emit_accept(const double p[3], const Box3d& bx) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (bx.hi[a] < p[a] || bx.lo[a] > p[a]) return false;  // outside: clamp() leaves a positive distance
    int best = -1;
    double best_d2 = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {  // +x,+y,+z through upperCorner, then -x,-y,-z through lowerCorner
        const int a = q % 3;
        double d[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double c = q < 3 ? bx.hi[b] : bx.lo[b];
            d[b] = (b == a) ? c - p[b] : ((p[b] - c) + c) - p[b];
        }
        double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        if (best < 0 || d2 < best_d2) { best = q; best_d2 = d2; }
    }
    const int a = best % 3;
    const double sgn = best < 3 ? 1.0 : -1.0;
    const double c = best < 3 ? bx.hi[a] : bx.lo[a];
    bool inside = (p[a] - c) * sgn < 0.0;
    return inside || best_d2 == 0.0;
}


// flags[t] = lattice point t is emitted; mine[t] (slab mode, else null) = ... and its cell plane belongs to this rank
__global__ void k_emit_flags(uint32_t* __restrict__ flags, uint32_t* __restrict__ mine, size_t total, unsigned cx, unsigned cy,
                             double spacing, double lx, double ly, double lz, Box3d bx, double oz, double ihz, int nz, int z0,
                             int z1) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > total) return;
    if (t == total) { flags[t] = 0; if (mine) mine[t] = 0; return; }
    double p[3];
    lattice_point(t, cx, cy, spacing, lx, ly, lz, p);
    const bool acc = box_emit_accept(p, bx);
    flags[t] = acc ? 1u : 0u;
    if (mine) {
        const int ck = cell_coord(p[2], oz, ihz, nz);
        mine[t] = (acc && ck >= z0 && ck < z1) ? 1u : 0u;
    }
}

// the same for an analytic implicit surface (one explicit surface or an ImplicitSurfaceSet3): signedDistance(p) <= 0
__global__ void k_emit_flags_shapes(uint32_t* __restrict__ flags, uint32_t* __restrict__ mine, size_t total, LatticeD L, ColliderD e,
                                    double oz, double ihz, int nz, int z0, int z1) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > total) return;
    if (t == total) { flags[t] = 0; if (mine) mine[t] = 0; return; }
    double p[3];
    lattice_point_l(L, t, p);
    const bool acc = cd_emit_accept(e, p);
    flags[t] = acc ? 1u : 0u;
    if (mine) {
        const int ck = cell_coord(p[2], oz, ihz, nz);
        mine[t] = (acc && ck >= z0 && ck < z1) ? 1u : 0u;
    }
}

// slot = exclusive scan of flags (position in the reference's emission order = the particle's original index),
// local = exclusive scan of mine (position among this rank's particles; == slot on a single GPU)
__global__ void k_emit_lattice(double* __restrict__ x, double* __restrict__ y, double* __restrict__ z, double* __restrict__ vx,
                               double* __restrict__ vy, double* __restrict__ vz, uint32_t* __restrict__ orig, size_t offset,
                               size_t orig_base, const uint32_t* __restrict__ slot, const uint32_t* __restrict__ local,
                               size_t total, unsigned cx, unsigned cy, double spacing, double lx, double ly, double lz,
                               double v0x, double v0y, double v0z) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    if (local[t + 1] == local[t]) return;  // rejected lattice point (or another rank's)
    double p[3];
    lattice_point(t, cx, cy, spacing, lx, ly, lz, p);
    size_t o = offset + local[t];
    x[o] = p[0]; y[o] = p[1]; z[o] = p[2];
    vx[o] = v0x; vy[o] = v0y; vz[o] = v0z;
    orig[o] = (uint32_t)(orig_base + slot[t]);
}

// the same for either point generator (LatticeD), used by the analytic-surface emitters
__global__ void k_emit_lattice_l(double* __restrict__ x, double* __restrict__ y, double* __restrict__ z, double* __restrict__ vx,
                                 double* __restrict__ vy, double* __restrict__ vz, uint32_t* __restrict__ orig, size_t offset,
                                 size_t orig_base, const uint32_t* __restrict__ slot, const uint32_t* __restrict__ local,
                                 size_t total, LatticeD L, double v0x, double v0y, double v0z) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    if (local[t + 1] == local[t]) return;
    double p[3];
    lattice_point_l(L, t, p);
    size_t o = offset + local[t];
    x[o] = p[0]; y[o] = p[1]; z[o] = p[2];
    vx[o] = v0x; vy[o] = v0y; vz[o] = v0z;
    orig[o] = (uint32_t)(orig_base + slot[t]);
}

// slab mode set_particles: keep the caller's particles whose cell plane this rank owns
__global__ void k_own_flags(const double* __restrict__ aos_pos, size_t n, double oz, double ihz, int nz, int z0, int z1,
                            uint32_t* __restrict__ mine) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { mine[i] = 0; return; }
    const int ck = cell_coord(aos_pos[3 * i + 2], oz, ihz, nz);
    mine[i] = (ck >= z0 && ck < z1) ? 1u : 0u;
}

__global__ void k_compact_aos3(const double* __restrict__ aos, const uint32_t* __restrict__ local, size_t n,
                               double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
                               uint32_t* __restrict__ orig, size_t id_base) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || local[i + 1] == local[i]) return;
    const size_t o = local[i];
    x[o] = aos[3 * i]; y[o] = aos[3 * i + 1]; z[o] = aos[3 * i + 2];
    if (orig) orig[o] = (uint32_t)(id_base + i);
}

// local (compact) download: positions/velocities in the sorted order plus their original indices
__global__ void k_soa_to_aos3(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                              double* __restrict__ aos, size_t n) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    aos[3 * s] = x[s]; aos[3 * s + 1] = y[s]; aos[3 * s + 2] = z[s];
}

__global__ void k_u32_to_u64(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__global__ void k_u64_to_u32(const uint64_t* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}

// ------------------------------------------------------------------------------------------------ collider
// Thin launch wrappers: the per-cell / per-face / per-particle bodies live in collider_ops.cuh as __host__ __device__
// functions, so that tests/hostcheck can run the very same source on the CPU against the oracle.
__global__ void __launch_bounds__(256) k_collider_sdf(GridD g, ColliderD c, double* __restrict__ sdf) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    sdf[((size_t)k * g.ny + j) * g.nx + i] = collider_sdf_cell(g, c, i, j, k);
}

template <int AXIS>
__global__ void __launch_bounds__(256) k_constrain_markers(GridD g, ColliderD c, const double* __restrict__ csdf,
                                                           double* __restrict__ data, char* __restrict__ marker) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx + (AXIS == 0) || j >= g.ny + (AXIS == 1) || k >= g.nz + (AXIS == 2)) return;
    constrain_marker_face<AXIS>(g, c, csdf, data, marker, i, j, k);
}

template <int AXIS>
__global__ void __launch_bounds__(256) k_boundary_velocity(GridD g, ColliderD c, double* __restrict__ bnd) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx + (AXIS == 0) || j >= g.ny + (AXIS == 1) || k >= g.nz + (AXIS == 2)) return;
    boundary_velocity_face<AXIS>(g, c, bnd, i, j, k);
}

template <int AXIS>
__global__ void __launch_bounds__(256) k_constrain_project(GridD g, ColliderD c, const double* __restrict__ csdf,
                                                           const double* __restrict__ u, const double* __restrict__ v,
                                                           const double* __restrict__ w, double* __restrict__ temp) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx + (AXIS == 0) || j >= g.ny + (AXIS == 1) || k >= g.nz + (AXIS == 2)) return;
    constrain_project_face<AXIS>(g, c, csdf, u, v, w, temp, i, j, k);
}

__global__ void __launch_bounds__(256) k_outside_collider_marker(GridD g, const double* __restrict__ csdf, char* __restrict__ marker) {
    int i = blockIdx.x * 64 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i >= g.nx || j >= g.ny) return;
    marker[((size_t)k * g.ny + j) * g.nx + i] = outside_collider_cell(g, csdf, i, j, k);
}

// col->resolveCollision(0.0, 0.0, &positions[i], &velocities[i]) — src/jet/pic_solver3.cpp:268-281
__global__ void k_resolve_collision(ColliderD c, AttrPtrs a, size_t np) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double pos[3] = {a.p[0][p], a.p[1][p], a.p[2][p]}, vel[3] = {a.p[3][p], a.p[4][p], a.p[5][p]};
    cd_resolve_collision(c, pos, vel);
    a.p[0][p] = pos[0]; a.p[1][p] = pos[1]; a.p[2][p] = pos[2];
    a.p[3][p] = vel[0]; a.p[4][p] = vel[1]; a.p[5][p] = vel[2];
}

// ================================================================================================ host side
static AttrPtrs attr_ptrs(jetcu_hybrid* h, int which, size_t offset = 0) {
    AttrPtrs a;
    for (int q = 0; q < NATTR; ++q) a.p[q] = h->attr[which][q] ? h->attr[which][q] + offset : nullptr;
    return a;
}
// the owned particles (everything on a single GPU; the middle block of the sorted array in slab mode)
static AttrPtrs owned_ptrs(jetcu_hybrid* h) { return attr_ptrs(h, h->cur, h->base); }

// Makes room for `need` particles per buffer; the owned block [base, base + np) of the current buffer survives and
// moves to offset 0.
static int hyb_reserve(jetcu_hybrid* h, uint64_t need) {
    jetcu_ctx* ctx = h->ctx;
    if (need <= h->cap) return JETCU_OK;
    uint64_t cap = need + need / 8 + 1024;
    for (int b = 0; b < 2; ++b) {
        for (int q = 0; q < h->nattr; ++q) {
            double* np_ = nullptr;
            JC_TRY(jc_alloc(ctx, &np_, cap));
            if (b == h->cur && h->np > 0) JC_TRY(jc_d2d(ctx, np_, h->attr[b][q] + h->base, h->np));
            if (q >= 6) JC_TRY(jc_zero(ctx, np_ + h->np, cap - h->np));
            jc_free(ctx, h->attr[b][q]);
            h->attr[b][q] = np_;
        }
        uint32_t* no = nullptr;
        JC_TRY(jc_alloc(ctx, &no, cap));
        if (b == h->cur && h->np > 0) JC_TRY(jc_d2d(ctx, no, h->orig[b] + h->base, h->np));
        jc_free(ctx, h->orig[b]);
        h->orig[b] = no;
    }
    uint32_t* t;
    jc_free(ctx, h->cell_of); jc_free(ctx, h->rank); jc_free(ctx, h->perm);
    JC_TRY(jc_alloc(ctx, &t, cap)); h->cell_of = t;
    JC_TRY(jc_alloc(ctx, &t, cap)); h->rank = t;
    JC_TRY(jc_alloc(ctx, &t, cap)); h->perm = t;
    h->cap = cap;
    h->base = 0;
    h->nsorted = h->np;
    h->binned = false; h->counted = false;
    return JETCU_OK;
}

// JETCU_BIN_TRACE=1: wall-clock of the binning's sub-stages with a stream synchronisation after each (diagnosis only: the
// synchronisations perturb the step), printed by every rank to stderr.
struct BinTrace {
    bool on;
    jetcu_ctx* ctx;
    double t0;
    char buf[512];
    int len = 0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit BinTrace(jetcu_ctx* c) : ctx(c) {
        static int env = -1;
        if (env < 0) { const char* e = getenv("JETCU_BIN_TRACE"); env = (e && atoi(e)) ? 1 : 0; }
        on = env != 0;
        if (on) { cudaStreamSynchronize(ctx->stream); t0 = now(); }
    }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(ctx->stream);
        const double t = now();
        len += snprintf(buf + len, sizeof(buf) - len, " %s %.3f", what, t - t0);
        t0 = t;
    }
    void flush(int rank) { if (on && len) fprintf(stderr, "[bin trace rank %d]%s\n", rank, buf); len = 0; }
};

// Slab mode, first half of the binning: copies every owned particle to the other ranks whose extended slab holds its
// (new) cell plane — migration and ghost refresh in one exchange — and appends what the other ranks sent to the owned
// block.  *ntot = owned + received.
static int hyb_slab_exchange(jetcu_hybrid* h, uint64_t* ntot) {
    jetcu_ctx* ctx = h->ctx;
    const int G = h->world, T = h->nz / G;
    const jetcu_grid& gg = h->P.grid;
    const size_t np = h->np;
    uint32_t* d_cnt = nullptr;   // [G] counts, [G] offsets, [G] fill cursors
    JC_TRY(jc_alloc(ctx, &d_cnt, 3 * (size_t)G));
    JC_TRY(jc_zero(ctx, d_cnt, 3 * (size_t)G));
    AttrPtrs own = owned_ptrs(h);
    uint32_t* orig = h->orig[h->cur] + h->base;
    BinTrace tr(ctx);
    if (np) JC_LAUNCH(ctx, k_slab_route<false>, jc_blocks(np, 256), 256, 0, own.p[2], np, gg.oz, 1.0 / gg.hz, h->nz, T, G, h->me,
                      d_cnt, nullptr, nullptr);
    uint32_t cnt32[64] = {};
    JC_TRY(jc_d2h(ctx, cnt32, d_cnt, (size_t)G));
    JC_TRY(jc_sync(ctx));
    tr.mark("route-count");
    uint64_t mine[64] = {}, all[64 * 64];
    for (int q = 0; q < G; ++q) mine[q] = cnt32[q];
    JC_TRY(comm_allgather_u64_host(ctx, mine, (size_t)G, all));
    tr.mark("allgather-counts");
    size_t send_off[65] = {}, recv_off[65] = {};
    uint32_t off32[64] = {};
    for (int q = 0; q < G; ++q) {
        send_off[q + 1] = send_off[q] + mine[q];
        recv_off[q + 1] = recv_off[q] + all[(size_t)q * G + h->me];
        off32[q] = (uint32_t)send_off[q];
    }
    const size_t nsend = send_off[G], nrecv = recv_off[G];
    const size_t rec = (size_t)h->nattr + 1;  // doubles per particle in a message
    uint32_t* d_list = nullptr;
    double *sendbuf = nullptr, *recvbuf = nullptr;
    int rc = JETCU_OK;
#define STEP(x) if (rc == JETCU_OK) rc = (x)
    STEP(jc_alloc(ctx, &d_list, nsend)); STEP(jc_alloc(ctx, &sendbuf, nsend * rec)); STEP(jc_alloc(ctx, &recvbuf, nrecv * rec));
    if (rc == JETCU_OK && nsend) {
        STEP(jc_h2d(ctx, d_cnt + G, off32, (size_t)G));
        if (rc == JETCU_OK) {
            k_slab_route<true><<<jc_blocks(np, 256), 256, 0, ctx->stream>>>(own.p[2], np, gg.oz, 1.0 / gg.hz, h->nz, T, G, h->me,
                                                                           d_cnt + 2 * G, d_cnt + G, d_list);
            ctx->launches++;
        }
        for (int q = 0; q < G && rc == JETCU_OK; ++q) {
            if (!mine[q]) continue;
            if (h->nattr == 6)
                k_slab_pack<6><<<jc_blocks(mine[q], 256), 256, 0, ctx->stream>>>(own, orig, d_list + send_off[q], (size_t)mine[q],
                                                                                sendbuf + send_off[q] * rec);
            else
                k_slab_pack<NATTR><<<jc_blocks(mine[q], 256), 256, 0, ctx->stream>>>(own, orig, d_list + send_off[q], (size_t)mine[q],
                                                                                    sendbuf + send_off[q] * rec);
            ctx->launches++;
        }
    }
    void *sp[64], *rp[64];
    size_t sb[64], rb[64];
    for (int q = 0; q < G; ++q) {
        sp[q] = sendbuf + send_off[q] * rec; sb[q] = (size_t)mine[q] * rec * sizeof(double);
        rp[q] = recvbuf + recv_off[q] * rec; rb[q] = (size_t)(recv_off[q + 1] - recv_off[q]) * rec * sizeof(double);
    }
    tr.mark("pack");
    STEP(comm_exchange(ctx, sp, sb, rp, rb));
    tr.mark("exchange");
    STEP(hyb_reserve(h, h->base + np + nrecv));  // may move the owned block to offset 0
    tr.mark("reserve");
    if (rc == JETCU_OK) {
        AttrPtrs dst = owned_ptrs(h);
        uint32_t* dorig = h->orig[h->cur] + h->base;
        for (int q = 0; q < G; ++q) {
            const size_t c = recv_off[q + 1] - recv_off[q];
            if (!c) continue;
            if (h->nattr == 6)
                k_slab_unpack<6><<<jc_blocks(c, 256), 256, 0, ctx->stream>>>(recvbuf + recv_off[q] * rec, c, dst, dorig, np + recv_off[q]);
            else
                k_slab_unpack<NATTR><<<jc_blocks(c, 256), 256, 0, ctx->stream>>>(recvbuf + recv_off[q] * rec, c, dst, dorig, np + recv_off[q]);
            ctx->launches++;
        }
    }
#undef STEP
    jc_free(ctx, d_list); jc_free(ctx, sendbuf); jc_free(ctx, recvbuf); jc_free(ctx, d_cnt);
    tr.mark("unpack");
    if (tr.on) { tr.len += snprintf(tr.buf + tr.len, sizeof(tr.buf) - tr.len, " | owned %zu sent %zu received %zu", np, nsend, nrecv); }
    tr.flush(h->me);
    *ntot = np + nrecv;
    return rc;
}

// Counting sort of the particles by owning cell; afterwards cell_start[c]..cell_start[c+1] is cell c's range.
// Slab mode: preceded by the migration / ghost exchange; the sorted array is [ghost-lo | owned | ghost-hi | dead].
static int hyb_bin(jetcu_hybrid* h) {
    jetcu_ctx* ctx = h->ctx;
    if (h->binned) return JETCU_OK;
    GridD g = gridd(h);
    const bool slab = h->world > 1;
    uint64_t ntot = h->np;
    if (slab) JC_TRY(hyb_slab_exchange(h, &ntot));
    BinTrace tr(ctx);
    const int e0 = slab ? max(h->z0 - SLAB_GHOST, 0) : 0, e1 = slab ? min(h->z1 + SLAB_GHOST, h->nz) : h->nz;
    // only the cells of the window [lo, hi) can receive particles (the rest go to the trash bin at index nc): count,
    // scan and per-cell ordering touch that range only, so the binning cost follows the slab, not the whole grid
    const size_t nxy = (size_t)h->nx * h->ny, lo = (size_t)e0 * nxy, hi = (size_t)e1 * nxy;
    // bins = (cell, octant) pairs, 8 per cell (k_cell_count); the trash bin sits behind the last real one
    const size_t blo = 8 * lo, bhi = 8 * hi, btrash = 8 * h->nc;
    if (!(h->counted && !slab)) {   // else: the last step's advection kernel already counted (count_cell)
        JC_TRY(jc_zero(ctx, h->sub_start + blo, bhi - blo + 1));
        JC_TRY(jc_zero(ctx, h->sub_start + btrash, 2));
        if (ntot > 0) {
            AttrPtrs src = owned_ptrs(h);
            JC_LAUNCH(ctx, k_cell_count, jc_blocks(ntot, 256), 256, 0, src.p[0], src.p[1], src.p[2], (size_t)ntot, g, e0, e1,
                      h->cell_of, h->rank, h->sub_start);
        }
    }
    h->counted = false;
    tr.mark("count");
    if (hi == h->nc) {
        JC_TRY(scan_exclusive_u32(ctx, h->sub_start + blo, h->sub_start + blo, bhi - blo + 2));
    } else {  // sub_start[bhi] = end of the window = where the trash bin starts
        JC_TRY(scan_exclusive_u32(ctx, h->sub_start + blo, h->sub_start + blo, bhi - blo + 1));
        JC_TRY(jc_d2d(ctx, h->sub_start + btrash, h->sub_start + bhi, 1));
    }
    tr.mark("scan");
    if (ntot > 0) JC_LAUNCH(ctx, k_scatter_perm, jc_blocks(ntot, 256), 256, 0, h->cell_of, h->rank, h->sub_start, h->perm, (size_t)ntot);
    // per-bin order by original index + the compact per-cell starts cell_start[lo .. hi]
    JC_LAUNCH(ctx, k_sort_cells, jc_blocks(bhi - blo + 1, 128), 128, 0, h->sub_start + blo, h->orig[h->cur] + h->base, h->perm,
              h->cell_start + lo, bhi - blo);
    JC_TRY(jc_d2d(ctx, h->cell_start + h->nc, h->sub_start + btrash, 2));
    if (ntot > 0) {
        if (h->nattr == 6)
            JC_LAUNCH(ctx, k_gather_attrs<6>, jc_blocks(ntot, 256), 256, 0, owned_ptrs(h), attr_ptrs(h, 1 - h->cur),
                      h->orig[h->cur] + h->base, h->orig[1 - h->cur], h->perm, (size_t)ntot);
        else
            JC_LAUNCH(ctx, k_gather_attrs<NATTR>, jc_blocks(ntot, 256), 256, 0, owned_ptrs(h), attr_ptrs(h, 1 - h->cur),
                      h->orig[h->cur] + h->base, h->orig[1 - h->cur], h->perm, (size_t)ntot);
        h->cur = 1 - h->cur;
    }
    tr.mark("perm-sort-gather");
    if (slab) {
        uint32_t marks[3] = {0, 0, 0};
        JC_TRY(jc_d2h(ctx, &marks[0], h->cell_start + (size_t)h->z0 * nxy, 1));
        JC_TRY(jc_d2h(ctx, &marks[1], h->cell_start + (size_t)h->z1 * nxy, 1));
        JC_TRY(jc_d2h(ctx, &marks[2], h->cell_start + h->nc, 1));
        JC_TRY(jc_sync(ctx));
        h->base = marks[0]; h->np = marks[1] - marks[0]; h->nsorted = marks[2];
        tr.mark("marks");
        if (tr.on) tr.len += snprintf(tr.buf + tr.len, sizeof(tr.buf) - tr.len, " | ntot %llu window cells %zu", (unsigned long long)ntot, hi - lo);
        tr.flush(h->me);
    } else {
        h->base = 0; h->nsorted = h->np;
    }
    h->binned = true;
    return JETCU_OK;
}

static int hyb_extrapolate_one(jetcu_hybrid* h, double* data, const char* marker, int sx, int sy, int sz, unsigned depth);

static inline dim3 face_grid(int sx, int sy, int sz) { return dim3((sx + 63) / 64, (sy + 3) / 4, sz); }

// The collider branch of GridFractionalBoundaryConditionSolver3::constrainVelocity
// (src/jet/grid_fractional_boundary_condition_solver3.cpp:55-187); grid-only work, replicated in slab mode.
static int hyb_constrain_collider(jetcu_hybrid* h) {
    jetcu_ctx* ctx = h->ctx;
    GridD g = gridd(h);
    const unsigned depth = (unsigned)ceil(h->P.max_cfl);  // GridFluidSolver3::applyBoundaryCondition, grid_fluid_solver3.cpp:346-353
    const dim3 blk(64, 4);
    const int nx = h->nx, ny = h->ny, nz = h->nz;
    JC_LAUNCH(ctx, k_constrain_markers<0>, face_grid(nx + 1, ny, nz), blk, 0, g, h->col, h->col_sdf, h->u, h->cm_u);
    JC_LAUNCH(ctx, k_constrain_markers<1>, face_grid(nx, ny + 1, nz), blk, 0, g, h->col, h->col_sdf, h->v, h->cm_v);
    JC_LAUNCH(ctx, k_constrain_markers<2>, face_grid(nx, ny, nz + 1), blk, 0, g, h->col, h->col_sdf, h->w, h->cm_w);
    JC_TRY(hyb_extrapolate_one(h, h->u, h->cm_u, nx + 1, ny, nz, depth));
    JC_TRY(hyb_extrapolate_one(h, h->v, h->cm_v, nx, ny + 1, nz, depth));
    JC_TRY(hyb_extrapolate_one(h, h->w, h->cm_w, nx, ny, nz + 1, depth));
    JC_LAUNCH(ctx, k_constrain_project<0>, face_grid(nx + 1, ny, nz), blk, 0, g, h->col, h->col_sdf, h->u, h->v, h->w, h->tmp_u);
    JC_LAUNCH(ctx, k_constrain_project<1>, face_grid(nx, ny + 1, nz), blk, 0, g, h->col, h->col_sdf, h->u, h->v, h->w, h->tmp_v);
    JC_LAUNCH(ctx, k_constrain_project<2>, face_grid(nx, ny, nz + 1), blk, 0, g, h->col, h->col_sdf, h->u, h->v, h->w, h->tmp_w);
    JC_TRY(jc_d2d(ctx, h->u, h->tmp_u, h->nu)); JC_TRY(jc_d2d(ctx, h->v, h->tmp_v, h->nv)); JC_TRY(jc_d2d(ctx, h->w, h->tmp_w, h->nw));
    return JETCU_OK;
}

// GridFluidSolver3::applyBoundaryCondition == constrainVelocity: the collider branch (if any), then the closed walls.
static int hyb_zero_walls(jetcu_hybrid* h) {
    jetcu_ctx* ctx = h->ctx;
    if (h->col.kind != 0) JC_TRY(hyb_constrain_collider(h));
    GridD g = gridd(h);
    size_t m = (size_t)g.ny * g.nz;
    if ((size_t)g.nx * g.nz > m) m = (size_t)g.nx * g.nz;
    if ((size_t)g.nx * g.ny > m) m = (size_t)g.nx * g.ny;
    JC_LAUNCH(ctx, k_zero_walls, jc_blocks(m, 256), 256, 0, g, h->P.closed_domain_boundary_flag, h->u, h->v, h->w);
    return JETCU_OK;
}

static int hyb_extrapolate_one(jetcu_hybrid* h, double* data, const char* marker, int sx, int sy, int sz, unsigned depth) {
    jetcu_ctx* ctx = h->ctx;
    size_t n = (size_t)sx * sy * sz;
    JC_TRY(jc_d2d(ctx, h->valid0, marker, n));
    char *v0 = h->valid0, *v1 = h->valid1;
    dim3 grid((sx + 63) / 64, (sy + 3) / 4, sz);
    for (unsigned it = 0; it < depth; ++it) {
        JC_LAUNCH(ctx, k_extrapolate_sweep, grid, dim3(64, 4), 0, data, v0, v1, sx, sy, sz);
        char* t = v0; v0 = v1; v1 = t;
    }
    return JETCU_OK;
}

// picks the sampler instantiation: exact-reciprocal normalisation when every spacing is a power of two, 32-bit sample
// indices when every component grid has fewer than 2^31 samples
#define HYB_SAMPLER_LAUNCH(kernel, grid, block, ...)                                                              \
    do {                                                                                                          \
        const bool pow2_ = g.ihx != 0.0 && g.ihy != 0.0 && g.ihz != 0.0;                                          \
        const bool i32_ = h->nu < (1ull << 31) && h->nv < (1ull << 31) && h->nw < (1ull << 31);                   \
        if (pow2_ && i32_) JC_LAUNCH(ctx, (kernel<true, int>), grid, block, 0, __VA_ARGS__);                      \
        else if (pow2_) JC_LAUNCH(ctx, (kernel<true, long long>), grid, block, 0, __VA_ARGS__);                   \
        else if (i32_) JC_LAUNCH(ctx, (kernel<false, int>), grid, block, 0, __VA_ARGS__);                         \
        else JC_LAUNCH(ctx, (kernel<false, long long>), grid, block, 0, __VA_ARGS__);                             \
    } while (0)

// advection kernels: <POW2, I, CC>; JETCU_MOVE_CACHE=0 selects the uncached form (A/B; results are bit-identical)
static inline int hyb_move_cache(int dflt) {
    static int on = -2;
    if (on == -2) { const char* e = getenv("JETCU_MOVE_CACHE"); on = e ? atoi(e) : -1; if (on < -1 || on > 5) on = -1; }
    return on < 0 ? dflt : on;
}
#define HYB_MOVE_LAUNCH_(kernel, CC_, grid, block, ...)                                                           \
    do {                                                                                                          \
        const bool pow2_ = g.ihx != 0.0 && g.ihy != 0.0 && g.ihz != 0.0;                                          \
        const bool i32_ = h->nu < (1ull << 31) && h->nv < (1ull << 31) && h->nw < (1ull << 31);                   \
        if (pow2_ && i32_) JC_LAUNCH(ctx, (kernel<true, int, CC_>), grid, block, 0, __VA_ARGS__);                 \
        else if (pow2_) JC_LAUNCH(ctx, (kernel<true, long long, CC_>), grid, block, 0, __VA_ARGS__);              \
        else if (i32_) JC_LAUNCH(ctx, (kernel<false, int, CC_>), grid, block, 0, __VA_ARGS__);                    \
        else JC_LAUNCH(ctx, (kernel<false, long long, CC_>), grid, block, 0, __VA_ARGS__);                        \
    } while (0)
// DFLT: the variant measured fastest for that kernel at 256^3 (GPU call 10, profiles/r02_move_variants.md)
#define HYB_MOVE_LAUNCH(kernel, DFLT, grid, block, ...)                                                           \
    do {                                                                                                          \
        const int cc_ = hyb_move_cache(DFLT);                                                                     \
        if (cc_ == 1) HYB_MOVE_LAUNCH_(kernel, 1, grid, block, __VA_ARGS__);                                      \
        else if (cc_ == 2) HYB_MOVE_LAUNCH_(kernel, 2, grid, block, __VA_ARGS__);                                 \
        else if (cc_ == 3) HYB_MOVE_LAUNCH_(kernel, 3, grid, block, __VA_ARGS__);                                 \
        else if (cc_ == 4) HYB_MOVE_LAUNCH_(kernel, 4, grid, block, __VA_ARGS__);                                 \
        else if (cc_ == 5) HYB_MOVE_LAUNCH_(kernel, 5, grid, block, __VA_ARGS__);                                 \
        else HYB_MOVE_LAUNCH_(kernel, 0, grid, block, __VA_ARGS__);                                               \
    } while (0)

#define HYB_ENTER(h)                                                      \
    if (!(h)) return JETCU_ERR_INVALID_ARG;                               \
    jetcu_ctx* ctx = (h)->ctx;                                            \
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");

extern "C" {

void jetcu_hybrid_default_params(jetcu_hybrid_params* p) {
    memset(p, 0, sizeof(*p));
    p->grid.nx = p->grid.ny = p->grid.nz = 1;
    p->grid.hx = p->grid.hy = p->grid.hz = 1.0;
    p->mode = JETCU_FLIP;
    p->closed_domain_boundary_flag = JETCU_DIR_ALL;
    p->gravity[1] = -9.8;
    p->max_cfl = 5.0;
    p->mgpcg.max_iterations = 100; p->mgpcg.max_levels = 5; p->mgpcg.n_restriction = 5; p->mgpcg.n_correction = 5;
    p->mgpcg.n_coarsest = 20; p->mgpcg.n_final = 20; p->mgpcg.tolerance = 1e-9; p->mgpcg.sor_factor = 1.5;
    p->mgpcg.use_red_black = 1;
}

int jetcu_hybrid_create(jetcu_ctx* ctx, const jetcu_hybrid_params* p, jetcu_hybrid** out) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");
    JC_ARG(ctx, p && out);
    JC_ARG(ctx, p->grid.nx > 0 && p->grid.ny > 0 && p->grid.nz > 0 && p->grid.nx < 32768 && p->grid.ny < 32768 && p->grid.nz < 32768);
    JC_ARG(ctx, p->grid.hx > 0 && p->grid.hy > 0 && p->grid.hz > 0);
    JC_ARG(ctx, p->mode >= JETCU_PIC && p->mode <= JETCU_APIC);
    // particle bins are (cell, octant) pairs held in 32-bit indices: up to 2^29 cells (812^3)
    if ((uint64_t)p->grid.nx * p->grid.ny * p->grid.nz * 8 >= (1ull << 32) - 2)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "the hybrid solver bins particles in 32-bit (cell, octant) indices: at most 2^29 cells");
    if (!p->mgpcg.use_red_black && jc_comm_world(ctx) > 1)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "lexicographic Gauss-Seidel (useRedBlackOrdering = false) is not slab-distributed: "
                          "use useRedBlackOrdering = true on a multi-GPU communicator");
    jetcu_hybrid* h = new jetcu_hybrid();
    h->ctx = ctx;
    h->P = *p;
    h->nx = (int)p->grid.nx; h->ny = (int)p->grid.ny; h->nz = (int)p->grid.nz;
    h->nu = (size_t)(h->nx + 1) * h->ny * h->nz;
    h->nv = (size_t)h->nx * (h->ny + 1) * h->nz;
    h->nw = (size_t)h->nx * h->ny * (h->nz + 1);
    h->nc = (size_t)h->nx * h->ny * h->nz;
    h->nattr = (p->mode == JETCU_APIC) ? NATTR : 6;
    h->world = jc_comm_world(ctx); h->me = jc_comm_rank(ctx);
    h->z0 = 0; h->z1 = h->nz;
    if (h->world > 1) {
        if (h->world > 64 || h->nz % (2 * h->world) != 0 || h->nz / h->world < 2 * SLAB_GHOST) {
            int rc0 = jetcu_fail(ctx, JETCU_ERR_INVALID_ARG, "z-slab mode needs nz (%d) to be a multiple of 2 x %d ranks with "
                                 "at least %d planes per rank", h->nz, h->world, 2 * SLAB_GHOST);
            delete h;
            return rc0;
        }
        h->z0 = h->me * (h->nz / h->world); h->z1 = (h->me + 1) * (h->nz / h->world);
    }
    size_t nmax = h->nu > h->nv ? h->nu : h->nv;
    if (h->nw > nmax) nmax = h->nw;
    int rc = JETCU_OK;
#define STEP(x) if (rc == JETCU_OK) rc = (x)
    STEP(jc_alloc(ctx, &h->u, h->nu)); STEP(jc_alloc(ctx, &h->v, h->nv)); STEP(jc_alloc(ctx, &h->w, h->nw));
    STEP(jc_zero(ctx, h->u, h->nu)); STEP(jc_zero(ctx, h->v, h->nv)); STEP(jc_zero(ctx, h->w, h->nw));
    STEP(jc_alloc(ctx, &h->um, h->nu)); STEP(jc_alloc(ctx, &h->vm, h->nv)); STEP(jc_alloc(ctx, &h->wm, h->nw));
    STEP(jc_zero(ctx, h->um, h->nu)); STEP(jc_zero(ctx, h->vm, h->nv)); STEP(jc_zero(ctx, h->wm, h->nw));
    STEP(jc_alloc(ctx, &h->valid0, nmax)); STEP(jc_alloc(ctx, &h->valid1, nmax));
    {
        const size_t nf[3] = {h->nu, h->nv, h->nw};
        for (int q = 0; q < 3; ++q) {
            STEP(jc_alloc(ctx, &h->wsum[q], nf[q])); STEP(jc_alloc(ctx, &h->pair[q], 2 * nf[q]));
        }
    }
    STEP(jc_alloc(ctx, &h->p2g_err, 1)); STEP(jc_zero(ctx, h->p2g_err, 1));
    STEP(jc_alloc(ctx, &h->sdf, h->nc));
    if (p->mode == JETCU_FLIP) {
        STEP(jc_alloc(ctx, &h->ud, h->nu)); STEP(jc_alloc(ctx, &h->vd, h->nv)); STEP(jc_alloc(ctx, &h->wd, h->nw));
        STEP(jc_zero(ctx, h->ud, h->nu)); STEP(jc_zero(ctx, h->vd, h->nv)); STEP(jc_zero(ctx, h->wd, h->nw));
    }
    STEP(jc_alloc(ctx, &h->cell_start, h->nc + 2));  // + the trash bin of slab mode + the end marker
    STEP(jc_zero(ctx, h->cell_start, h->nc + 2));
    STEP(jc_alloc(ctx, &h->sub_start, 8 * h->nc + 2));
    STEP(jc_zero(ctx, h->sub_start, 8 * h->nc + 2));
    if (rc == JETCU_OK) {
        // signedDistanceField is created with kMaxD (src/jet/pic_solver3.cpp:29-30)
        k_fill_f64<<<jc_blocks(h->nc, 256), 256, 0, ctx->stream>>>(h->sdf, DBL_MAX, h->nc);
        ctx->launches++;
    }
    STEP(pressure_prepare(ctx, &h->ps, &p->grid, p->mgpcg.max_levels));
    STEP(jc_zero(ctx, h->ps.x0, h->nc));
    for (int e = 0; e < 12 && rc == JETCU_OK; ++e)
        if (cudaEventCreate(&h->ev[e]) != cudaSuccess) rc = jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaEventCreate failed");
    STEP(jc_sync(ctx));
#undef STEP
    if (rc != JETCU_OK) { jetcu_hybrid_destroy(h); return rc; }
    *out = h;
    return JETCU_OK;
}

int jetcu_hybrid_update_params(jetcu_hybrid* h, const jetcu_hybrid_params* p) {
    HYB_ENTER(h);
    JC_ARG(ctx, p != nullptr);
    if (memcmp(&p->grid, &h->P.grid, sizeof(p->grid)) != 0 || p->mode != h->P.mode)
        return jetcu_fail(ctx, JETCU_ERR_INVALID_ARG, "jetcu_hybrid_update_params: grid and mode are fixed at creation");
    JC_ARG(ctx, p->max_cfl > 0.0);
    h->P = *p;   // gravity, maxCfl, picBlendingFactor, closedDomainBoundaryFlag, MGPCG parameters: read at every stage call
    return JETCU_OK;
}

int jetcu_hybrid_destroy(jetcu_hybrid* h) {
    if (!h) return JETCU_OK;
    jetcu_ctx* ctx = h->ctx;
    cudaSetDevice(ctx->device);
    jc_free(ctx, h->u); jc_free(ctx, h->v); jc_free(ctx, h->w);
    jc_free(ctx, h->um); jc_free(ctx, h->vm); jc_free(ctx, h->wm);
    jc_free(ctx, h->valid0); jc_free(ctx, h->valid1); jc_free(ctx, h->sdf);
    jc_free(ctx, h->ud); jc_free(ctx, h->vd); jc_free(ctx, h->wd);
    for (int q = 0; q < 3; ++q) { jc_free(ctx, h->wsum[q]); jc_free(ctx, h->pair[q]); }
    jc_free(ctx, h->p2g_err);
    jc_free(ctx, h->col_sdf); jc_free(ctx, h->bnd_u); jc_free(ctx, h->bnd_v); jc_free(ctx, h->bnd_w);
    jc_free(ctx, h->tmp_u); jc_free(ctx, h->tmp_v); jc_free(ctx, h->tmp_w);
    jc_free(ctx, h->cm_u); jc_free(ctx, h->cm_v); jc_free(ctx, h->cm_w);
    for (int b = 0; b < 2; ++b) {
        for (int q = 0; q < NATTR; ++q) jc_free(ctx, h->attr[b][q]);
        jc_free(ctx, h->orig[b]);
    }
    jc_free(ctx, h->cell_of); jc_free(ctx, h->rank); jc_free(ctx, h->perm); jc_free(ctx, h->cell_start); jc_free(ctx, h->sub_start);
    pressure_free(ctx, &h->ps);
    for (int e = 0; e < 12; ++e) if (h->ev[e]) cudaEventDestroy(h->ev[e]);
    cudaStreamSynchronize(ctx->stream);
    delete h;
    return JETCU_OK;
}

// Uploads n AoS particles; on a single GPU all of them, in slab mode (own_only) those whose cell plane this rank owns.
// ids == nullptr: original index = position in the caller's array (+ id_base).
// append: the uploaded particles join the device-resident set (which keeps its state, APIC affine vectors included) as
// original indices nglobal .. nglobal + n - 1; otherwise they replace it.
static int hyb_upload(jetcu_hybrid* h, uint64_t n, const double* positions, const double* velocities, const uint64_t* ids,
                      bool own_only, bool append = false) {
    jetcu_ctx* ctx = h->ctx;
    const jetcu_grid& gg = h->P.grid;
    if (!append) { h->np = 0; h->base = 0; h->nsorted = 0; }
    h->binned = false; h->counted = false;
    if (n == 0) return JETCU_OK;
    double* stage = nullptr;
    uint32_t* local = nullptr;
    uint64_t* dids = nullptr;
    int rc = jc_alloc(ctx, &stage, 3 * n);
#define STEP(x) if (rc == JETCU_OK) rc = (x)
    STEP(jc_alloc(ctx, &local, n + 1));
    STEP(jc_h2d(ctx, stage, positions, 3 * n));
    uint64_t keep = n;
    if (rc == JETCU_OK) {
        // own_only: flag by cell plane; otherwise every particle is kept (flag 1 for i < n)
        const int z0 = own_only ? h->z0 : 0, z1 = own_only ? h->z1 : h->nz;
        k_own_flags<<<jc_blocks(n + 1, 256), 256, 0, ctx->stream>>>(stage, (size_t)n, gg.oz, 1.0 / gg.hz, h->nz, z0, z1, local);
        ctx->launches++;
        rc = scan_exclusive_u32(ctx, local, local, n + 1);
        uint32_t k32 = 0;
        STEP(jc_d2h(ctx, &k32, local + n, 1));
        STEP(jc_sync(ctx));
        keep = k32;
    }
    STEP(hyb_reserve(h, (append ? h->base + h->np : 0) + keep));
    const int c = h->cur;
    const size_t at = append ? h->base + h->np : 0;   // after hyb_reserve: growing moves the owned block to offset 0
    if (rc == JETCU_OK && keep > 0) {
        k_compact_aos3<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(stage, local, (size_t)n, h->attr[c][0] + at, h->attr[c][1] + at,
                                                                  h->attr[c][2] + at, ids ? nullptr : h->orig[c] + at,
                                                                  append ? (size_t)h->nglobal : 0);
        ctx->launches++;
        if (velocities) {
            STEP(jc_h2d(ctx, stage, velocities, 3 * n));
            if (rc == JETCU_OK) {
                k_compact_aos3<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(stage, local, (size_t)n, h->attr[c][3] + at,
                                                                          h->attr[c][4] + at, h->attr[c][5] + at, nullptr, 0);
                ctx->launches++;
            }
        } else {
            for (int q = 3; q < 6 && rc == JETCU_OK; ++q) rc = jc_zero(ctx, h->attr[c][q] + at, keep);
        }
        // new particles start with zero affine vectors: _cX.resize(n) value-initialises the tail (apic_solver3.cpp:36-38)
        for (int q = 6; q < h->nattr && rc == JETCU_OK; ++q) rc = jc_zero(ctx, h->attr[c][q] + at, keep);
        if (ids) {  // caller-supplied original indices (every particle kept: own_only is false on this path)
            STEP(jc_alloc(ctx, &dids, n));
            STEP(jc_h2d(ctx, dids, ids, n));
            if (rc == JETCU_OK) {
                k_u64_to_u32<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(dids, h->orig[c] + at, (size_t)n);
                ctx->launches++;
            }
        }
    }
    if (rc == JETCU_OK) { h->np = (append ? h->np : 0) + keep; h->nsorted = h->np; rc = jc_sync(ctx); }
#undef STEP
    jc_free(ctx, stage); jc_free(ctx, local); jc_free(ctx, dids);
    return rc;
}

int jetcu_hybrid_set_particles(jetcu_hybrid* h, uint64_t n, const double* positions, const double* velocities) {
    HYB_ENTER(h);
    JC_ARG(ctx, n == 0 || positions != nullptr);
    JC_ARG(ctx, n < (1ull << 32) - 1);
    JC_TRY(hyb_upload(h, n, positions, velocities, nullptr, h->world > 1));
    h->nglobal = n;
    return JETCU_OK;
}

int jetcu_hybrid_append_particles(jetcu_hybrid* h, uint64_t n_new, const double* positions, const double* velocities) {
    HYB_ENTER(h);
    JC_ARG(ctx, n_new == 0 || positions != nullptr);
    JC_ARG(ctx, h->nglobal + n_new < (1ull << 32) - 1);
    JC_TRY(hyb_upload(h, n_new, positions, velocities, nullptr, h->world > 1, true));
    h->nglobal += n_new;
    return JETCU_OK;
}

int jetcu_hybrid_set_local_particles(jetcu_hybrid* h, uint64_t n_local, const uint64_t* ids, const double* positions,
                                     const double* velocities, uint64_t n_global) {
    HYB_ENTER(h);
    JC_ARG(ctx, n_local == 0 || (positions != nullptr && ids != nullptr));
    JC_ARG(ctx, n_global < (1ull << 32) - 1 && n_local <= n_global);
    JC_TRY(hyb_upload(h, n_local, positions, velocities, ids, false));
    h->nglobal = n_global;
    return JETCU_OK;
}

// Emission over the lattice of region [lo, hi] (GridPointGenerator3::forEachPoint, src/jet/grid_point_generator3.cpp:15-38):
// flag kernel (box fast path or analytic surface) -> exclusive scan = the reference's emission order -> compaction.
static int hyb_emit_lattice(jetcu_hybrid* h, const double lo[3], const double hi[3], double spacing, const double v0[3],
                            uint64_t* n_emitted, const Box3d* box, const ColliderD* surf, int bcc = 0) {
    jetcu_ctx* ctx = h->ctx;
    const jetcu_grid& g = h->P.grid;
    LatticeD L;
    const uint64_t total = cd_lattice_make(lo, hi, spacing, bcc, &L);   // the generators' own loop tests
    const unsigned cnt[3] = {L.cx0, L.cy0, L.cz};                       // the grid lattice of the Box3 fast path
    JC_ARG(ctx, h->nglobal + total < (1ull << 32) - 2);
    const bool slab = h->world > 1;
    uint64_t accepted = 0, accepted_here = 0;
    if (total > 0) {
        uint32_t *slot = nullptr, *mine = nullptr;
        JC_TRY(jc_alloc(ctx, &slot, total + 1));
        int rc = JETCU_OK;
        if (slab) rc = jc_alloc(ctx, &mine, total + 1);
        if (rc == JETCU_OK) {
            if (box)
                k_emit_flags<<<jc_blocks(total + 1, 256), 256, 0, ctx->stream>>>(slot, mine, (size_t)total, cnt[0], cnt[1], spacing,
                                                                                  lo[0], lo[1], lo[2], *box, g.oz, 1.0 / g.hz, h->nz,
                                                                                  h->z0, h->z1);
            else
                k_emit_flags_shapes<<<jc_blocks(total + 1, 256), 256, 0, ctx->stream>>>(slot, mine, (size_t)total, L, *surf, g.oz,
                                                                                         1.0 / g.hz, h->nz, h->z0, h->z1);
            ctx->launches++;
            rc = scan_exclusive_u32(ctx, slot, slot, total + 1);
        }
        if (rc == JETCU_OK && slab) rc = scan_exclusive_u32(ctx, mine, mine, total + 1);
        uint32_t n_acc = 0, n_mine = 0;
        if (rc == JETCU_OK) rc = jc_d2h(ctx, &n_acc, slot + total, 1);
        if (rc == JETCU_OK && slab) rc = jc_d2h(ctx, &n_mine, mine + total, 1);
        if (rc == JETCU_OK) rc = jc_sync(ctx);
        accepted = n_acc;
        accepted_here = slab ? n_mine : n_acc;
        if (rc == JETCU_OK) rc = hyb_reserve(h, h->base + h->np + accepted_here);
        if (rc == JETCU_OK && accepted_here > 0) {
            int c = h->cur;
            double z0[3] = {0, 0, 0};
            const double* vv = v0 ? v0 : z0;
            const size_t at = h->base + h->np;
            if (box)
                k_emit_lattice<<<jc_blocks(total, 256), 256, 0, ctx->stream>>>(
                    h->attr[c][0], h->attr[c][1], h->attr[c][2], h->attr[c][3], h->attr[c][4], h->attr[c][5], h->orig[c], at,
                    (size_t)h->nglobal, slot, slab ? mine : slot, (size_t)total, cnt[0], cnt[1], spacing, lo[0], lo[1], lo[2], vv[0],
                    vv[1], vv[2]);
            else
                k_emit_lattice_l<<<jc_blocks(total, 256), 256, 0, ctx->stream>>>(
                    h->attr[c][0], h->attr[c][1], h->attr[c][2], h->attr[c][3], h->attr[c][4], h->attr[c][5], h->orig[c], at,
                    (size_t)h->nglobal, slot, slab ? mine : slot, (size_t)total, L, vv[0], vv[1], vv[2]);
            ctx->launches++;
            for (int q = 6; q < h->nattr && rc == JETCU_OK; ++q) rc = jc_zero(ctx, h->attr[c][q] + at, accepted_here);
        }
        jc_free(ctx, slot); jc_free(ctx, mine);
        if (rc != JETCU_OK) return rc;
    }
    h->np += accepted_here;
    h->nglobal += accepted;
    h->nsorted = h->np;  // ghosts (if any) are stale now; the next binning rebuilds them
    h->binned = false; h->counted = false;
    if (n_emitted) *n_emitted = accepted;
    return jc_sync(ctx);
}

int jetcu_hybrid_emit_box(jetcu_hybrid* h, const double lower[3], const double upper[3], double spacing,
                          const double v0[3], uint64_t* n_emitted) {
    HYB_ENTER(h);
    JC_ARG(ctx, lower && upper && spacing > 0);
    const jetcu_grid& g = h->P.grid;
    // region = box ∩ domain bounding box (VolumeParticleEmitter3::emit, src/jet/volume_particle_emitter3.cpp:74-79)
    double dlo[3] = {g.ox, g.oy, g.oz};
    double dhi[3] = {g.ox + g.hx * (double)g.nx, g.oy + g.hy * (double)g.ny, g.oz + g.hz * (double)g.nz};
    double lo[3], hi[3];
    Box3d bx;
    for (int a = 0; a < 3; ++a) {
        // ImplicitSurface3::boundingBox() = Transform3::toWorld applied twice (Box3's and SurfaceToImplicit3's identity
        // transforms: src/jet/surface3.cpp:27-29, surface_to_implicit3.cpp:63-65); each pass rebuilds the corners as
        // size * (-+1/2) + midPoint (include/jet/detail/bounding_box3-inl.h:150-152,201-208), which is not exact.
        double l = lower[a], u = upper[a];
        for (int pass = 0; pass < 2; ++pass) {
            double w = u - l, mid = (u + l) / 2.0;
            double c0 = w * -0.5 + mid, c1 = w * 0.5 + mid;
            l = fmin(c0, c1); u = fmax(c0, c1);
        }
        lo[a] = fmax(dlo[a], l); hi[a] = fmin(dhi[a], u);
        bx.lo[a] = lower[a]; bx.hi[a] = upper[a];
    }
    return hyb_emit_lattice(h, lo, hi, spacing, v0, n_emitted, &bx, nullptr);
}

// VolumeParticleEmitter3 over an analytic implicit surface: the emitters of the hybrid_liquid_sim examples —
// half-space + sphere (main.cpp:110-141), a set of two boxes (:220-246) — on the device, in the reference's order.
int jetcu_hybrid_emit_shapes(jetcu_hybrid* h, const jetcu_collider_shape* shapes, uint32_t n, int as_set, int lattice, double spacing,
                             const double v0[3], uint64_t* n_emitted) {
    HYB_ENTER(h);
    JC_ARG(ctx, shapes && n >= 1 && n <= JC_COLLIDER_MAX_SHAPES && spacing > 0 && (n == 1 || as_set));
    JC_ARG(ctx, lattice == JETCU_LATTICE_GRID || lattice == JETCU_LATTICE_BCC);
    for (uint32_t i = 0; i < n; ++i) {
        JC_ARG(ctx, shapes[i].kind >= JETCU_COLLIDER_BOX && shapes[i].kind <= JETCU_COLLIDER_PLANE);
        if (shapes[i].kind == JETCU_COLLIDER_SPHERE || shapes[i].kind == JETCU_COLLIDER_CYLINDER) JC_ARG(ctx, shapes[i].b[0] > 0.0);
    }
    const ColliderD e = cd_make_collider(shapes, n, as_set, nullptr, nullptr, 0.0);
    const jetcu_grid& g = h->P.grid;
    double lo[3] = {g.ox, g.oy, g.oz};
    double hi[3] = {g.ox + g.hx * (double)g.nx, g.oy + g.hy * (double)g.ny, g.oz + g.hz * (double)g.nz};
    cd_emit_region(e, lo, hi);
    return hyb_emit_lattice(h, lo, hi, spacing, v0, n_emitted, nullptr, &e, lattice == JETCU_LATTICE_BCC ? 1 : 0);
}

uint64_t jetcu_hybrid_num_particles(const jetcu_hybrid* h) { return h ? h->nglobal : 0; }
uint64_t jetcu_hybrid_num_local_particles(const jetcu_hybrid* h) { return h ? h->np : 0; }

static int download_vec3(jetcu_hybrid* h, int a0, double* host) {
    jetcu_ctx* ctx = h->ctx;
    if (!host || h->nglobal == 0) return JETCU_OK;
    double* stage = nullptr;
    JC_TRY(jc_alloc(ctx, &stage, 3 * h->nglobal));
    // slab mode: entries of particles owned by other ranks stay zero, so summing the ranks' arrays gives the whole set
    if (h->world > 1) JC_TRY(jc_zero(ctx, stage, 3 * h->nglobal));
    int c = h->cur;
    if (h->np) {
        k_soa_to_aos3_orig<<<jc_blocks(h->np, 256), 256, 0, ctx->stream>>>(h->attr[c][a0] + h->base, h->attr[c][a0 + 1] + h->base,
                                                                          h->attr[c][a0 + 2] + h->base, h->orig[c] + h->base, stage,
                                                                          h->np);
        ctx->launches++;
    }
    int rc = jc_d2h(ctx, host, stage, 3 * h->nglobal);
    if (rc == JETCU_OK) rc = jc_sync(ctx);
    jc_free(ctx, stage);
    return rc;
}

int jetcu_hybrid_get_particles(jetcu_hybrid* h, double* positions, double* velocities) {
    HYB_ENTER(h);
    JC_TRY(download_vec3(h, 0, positions));
    return download_vec3(h, 3, velocities);
}

int jetcu_hybrid_get_local_particles(jetcu_hybrid* h, uint64_t* ids, double* positions, double* velocities) {
    HYB_ENTER(h);
    const size_t np = h->np;
    if (np == 0) return JETCU_OK;
    const int c = h->cur;
    double* stage = nullptr;
    JC_TRY(jc_alloc(ctx, &stage, 3 * np));
    int rc = JETCU_OK;
    for (int part = 0; part < 2 && rc == JETCU_OK; ++part) {
        double* host = part == 0 ? positions : velocities;
        if (!host) continue;
        const int a0 = 3 * part;
        k_soa_to_aos3<<<jc_blocks(np, 256), 256, 0, ctx->stream>>>(h->attr[c][a0] + h->base, h->attr[c][a0 + 1] + h->base,
                                                                  h->attr[c][a0 + 2] + h->base, stage, np);
        ctx->launches++;
        rc = jc_d2h(ctx, host, stage, 3 * np);
    }
    if (rc == JETCU_OK && ids) {
        uint64_t* wide = reinterpret_cast<uint64_t*>(stage);
        k_u32_to_u64<<<jc_blocks(np, 256), 256, 0, ctx->stream>>>(h->orig[c] + h->base, wide, np);
        ctx->launches++;
        rc = jc_d2h(ctx, ids, wide, np);
    }
    if (rc == JETCU_OK) rc = jc_sync(ctx);
    jc_free(ctx, stage);
    return rc;
}

int jetcu_hybrid_get_affine(jetcu_hybrid* h, double* cx, double* cy, double* cz) {
    HYB_ENTER(h);
    if (h->P.mode != JETCU_APIC) return jetcu_fail(ctx, JETCU_ERR_INVALID_ARG, "affine vectors exist only in APIC mode");
    JC_TRY(download_vec3(h, 6, cx));
    JC_TRY(download_vec3(h, 9, cy));
    return download_vec3(h, 12, cz);
}

int jetcu_hybrid_set_collider(jetcu_hybrid* h, const double* collider_sdf, const double*, const double*, const double*,
                              double) {
    HYB_ENTER(h);
    if (collider_sdf)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED,
                          "a collider given only as an SDF grid cannot resolve particle collisions (the reference "
                          "queries the analytic surface): use jetcu_hybrid_set_collider_shape");
    return JETCU_OK;
}

int jetcu_hybrid_set_collider_shapes(jetcu_hybrid* h, const jetcu_collider_shape* shapes, uint32_t n, int as_set,
                                     const double linear_velocity[3], const double angular_velocity[3], double friction) {
    HYB_ENTER(h);
    JC_ARG(ctx, n <= JC_COLLIDER_MAX_SHAPES && (n == 0 || shapes != nullptr));
    JC_ARG(ctx, n <= 1 || as_set);  // several surfaces exist in Jet only as members of an ImplicitSurfaceSet3
    for (uint32_t i = 0; i < n; ++i) {
        JC_ARG(ctx, shapes[i].kind >= JETCU_COLLIDER_BOX && shapes[i].kind <= JETCU_COLLIDER_PLANE);
        if (shapes[i].kind == JETCU_COLLIDER_SPHERE || shapes[i].kind == JETCU_COLLIDER_CYLINDER) JC_ARG(ctx, shapes[i].b[0] > 0.0);
    }
    h->col = ColliderD();
    if (n == 0) return JETCU_OK;
    const ColliderD c = cd_make_collider(shapes, n, as_set, linear_velocity, angular_velocity, friction);
    if (!h->col_sdf) {
        JC_TRY(jc_alloc(ctx, &h->col_sdf, h->nc));
        JC_TRY(jc_alloc(ctx, &h->bnd_u, h->nu)); JC_TRY(jc_alloc(ctx, &h->bnd_v, h->nv)); JC_TRY(jc_alloc(ctx, &h->bnd_w, h->nw));
        JC_TRY(jc_alloc(ctx, &h->tmp_u, h->nu)); JC_TRY(jc_alloc(ctx, &h->tmp_v, h->nv)); JC_TRY(jc_alloc(ctx, &h->tmp_w, h->nw));
        JC_TRY(jc_alloc(ctx, &h->cm_u, h->nu)); JC_TRY(jc_alloc(ctx, &h->cm_v, h->nv)); JC_TRY(jc_alloc(ctx, &h->cm_w, h->nw));
    }
    GridD g = gridd(h);
    const dim3 blk(64, 4);
    // the reference re-rasterises the (static) shape at every beginAdvanceTimeStep (grid_fluid_solver3.cpp:448-453):
    // a pure function of the shape, so once here
    JC_LAUNCH(ctx, k_collider_sdf, face_grid(h->nx, h->ny, h->nz), blk, 0, g, c, h->col_sdf);
    JC_LAUNCH(ctx, k_boundary_velocity<0>, face_grid(h->nx + 1, h->ny, h->nz), blk, 0, g, c, h->bnd_u);
    JC_LAUNCH(ctx, k_boundary_velocity<1>, face_grid(h->nx, h->ny + 1, h->nz), blk, 0, g, c, h->bnd_v);
    JC_LAUNCH(ctx, k_boundary_velocity<2>, face_grid(h->nx, h->ny, h->nz + 1), blk, 0, g, c, h->bnd_w);
    JC_TRY(jc_sync(ctx));
    h->col = c;
    return JETCU_OK;
}

int jetcu_hybrid_set_collider_shape(jetcu_hybrid* h, int kind, const double a[3], const double b[3],
                                    const double linear_velocity[3], const double angular_velocity[3], double friction) {
    HYB_ENTER(h);
    JC_ARG(ctx, kind >= JETCU_COLLIDER_NONE && kind <= JETCU_COLLIDER_PLANE);
    if (kind == JETCU_COLLIDER_NONE) return jetcu_hybrid_set_collider_shapes(h, nullptr, 0, 0, nullptr, nullptr, 0.0);
    JC_ARG(ctx, a != nullptr && b != nullptr);
    jetcu_collider_shape s;
    memset(&s, 0, sizeof(s));
    s.kind = kind;
    for (int q = 0; q < 3; ++q) { s.a[q] = a[q]; s.b[q] = b[q]; }
    return jetcu_hybrid_set_collider_shapes(h, &s, 1, 0, linear_velocity, angular_velocity, friction);
}

int jetcu_hybrid_get_collider_sdf(jetcu_hybrid* h, double* sdf) {
    HYB_ENTER(h);
    JC_ARG(ctx, sdf);
    if (h->col.kind == 0) { for (size_t i = 0; i < h->nc; ++i) sdf[i] = DBL_MAX; return JETCU_OK; }  // fill(kMaxD), :272
    JC_TRY(jc_d2h(ctx, sdf, h->col_sdf, h->nc));
    return jc_sync(ctx);
}

int jetcu_hybrid_set_velocity(jetcu_hybrid* h, const double* u, const double* v, const double* w) {
    HYB_ENTER(h);
    if (u) JC_TRY(jc_h2d(ctx, h->u, u, h->nu));
    if (v) JC_TRY(jc_h2d(ctx, h->v, v, h->nv));
    if (w) JC_TRY(jc_h2d(ctx, h->w, w, h->nw));
    return jc_sync(ctx);
}

int jetcu_hybrid_get_velocity(jetcu_hybrid* h, double* u, double* v, double* w) {
    HYB_ENTER(h);
    if (u) JC_TRY(jc_d2h(ctx, u, h->u, h->nu));
    if (v) JC_TRY(jc_d2h(ctx, v, h->v, h->nv));
    if (w) JC_TRY(jc_d2h(ctx, w, h->w, h->nw));
    return jc_sync(ctx);
}

int jetcu_hybrid_get_sdf(jetcu_hybrid* h, double* sdf) {
    HYB_ENTER(h);
    JC_ARG(ctx, sdf);
    JC_TRY(jc_d2h(ctx, sdf, h->sdf, h->nc));
    return jc_sync(ctx);
}

int jetcu_hybrid_get_pressure(jetcu_hybrid* h, double* p) {
    HYB_ENTER(h);
    JC_ARG(ctx, p);
    JC_TRY(jc_d2h(ctx, p, h->ps.x0, h->nc));
    return jc_sync(ctx);
}

int jetcu_hybrid_get_markers(jetcu_hybrid* h, char* um, char* vm, char* wm) {
    HYB_ENTER(h);
    if (um) JC_TRY(jc_d2h(ctx, um, h->um, h->nu));
    if (vm) JC_TRY(jc_d2h(ctx, vm, h->vm, h->nv));
    if (wm) JC_TRY(jc_d2h(ctx, wm, h->wm, h->nw));
    return jc_sync(ctx);
}

int jetcu_hybrid_cfl(jetcu_hybrid* h, double dt, double* cfl) {
    HYB_ENTER(h);
    JC_ARG(ctx, cfl);
    GridD g = gridd(h);
    const double* gr = h->P.gravity;
    JC_LAUNCH(ctx, k_cfl_partial, CFL_GRID, CFL_BLOCK, 0, g, h->u, h->v, h->w, dt * gr[0], dt * gr[1], dt * gr[2],
              ctx->d_partials);
    JC_LAUNCH(ctx, k_max_final, 1, CFL_BLOCK, 0, ctx->d_partials, CFL_GRID, ctx->d_scalars);
    JC_TRY(jc_d2h(ctx, ctx->h_scalars, ctx->d_scalars, 1));
    JC_TRY(jc_sync(ctx));
    double minH = fmin(fmin(g.hx, g.hy), g.hz);
    *cfl = ctx->h_scalars[0] * dt / minH;
    return JETCU_OK;
}

int jetcu_hybrid_num_sub_steps(jetcu_hybrid* h, double dt, uint32_t* n) {
    HYB_ENTER(h);
    JC_ARG(ctx, n);
    double c;
    JC_TRY(jetcu_hybrid_cfl(h, dt, &c));
    *n = (uint32_t)fmax(ceil(c / h->P.max_cfl), 1.0);  // src/jet/grid_fluid_solver3.cpp:212-217
    return JETCU_OK;
}

int jetcu_hybrid_apply_boundary_condition(jetcu_hybrid* h) {
    HYB_ENTER(h);
    return hyb_zero_walls(h);
}

// Slab mode: this rank's planes of a face field become valid on every rank.  Cell-plane-partitioned part by an
// in-place all-gather; the w-type field has one more plane (k = nz), owned by the last rank and broadcast.
static int hyb_gather_planes(jetcu_hybrid* h, void* field, size_t plane_bytes, bool extra_top_plane) {
    jetcu_ctx* ctx = h->ctx;
    if (h->world <= 1) return JETCU_OK;
    JC_TRY(comm_allgather_inplace(ctx, field, (size_t)(h->z1 - h->z0) * plane_bytes));
    if (extra_top_plane) JC_TRY(comm_bcast(ctx, static_cast<char*>(field) + (size_t)h->nz * plane_bytes, plane_bytes, h->world - 1));
    return JETCU_OK;
}

int jetcu_hybrid_transfer_particles_to_grids(jetcu_hybrid* h) {
    HYB_ENTER(h);
    JC_TRY(hyb_bin(h));
    GridD g = gridd(h);
    AttrPtrs a = attr_ptrs(h, h->cur);  // whole sorted array: cell_start indexes it from 0 (ghosts included)
    const bool apic = h->P.mode == JETCU_APIC;
    const bool slab = h->world > 1;
    int slot = jc_prof_begin(ctx, JETCU_PROF_P2G);
    const bool flip = h->P.mode == JETCU_FLIP;
    // planes of each face field this rank finalises: all of them, or (slab mode) the ones it owns — their sums are
    // complete because the ghost particles of the neighbouring planes were summed along, cell by cell and tile by tile in
    // the same order as on a single GPU (p2g_ops.cuh)
    const size_t pu = (size_t)(h->nx + 1) * h->ny, pv = (size_t)h->nx * (h->ny + 1), pw = (size_t)h->nx * h->ny;
    const size_t k0 = (size_t)h->z0, nk = (size_t)(h->z1 - h->z0);
    const size_t nkw = nk + ((!slab || h->me == h->world - 1) ? 1 : 0);
    const int kb = slab ? max(h->z0 - SLAB_GHOST, 0) : 0, ke = slab ? min(h->z1 + SLAB_GHOST, h->nz) : h->nz;
    {   // u, v, w double as the value sums; weight sums and markers start from zero as well.  Only the planes the cells of
        // this rank can reach (its particle window + one ring) need it; the others are overwritten by the all-gather below
        const size_t zlo = (size_t)max(kb - 1, 0), zhi = (size_t)min(ke + 1, h->nz);
        const size_t pl[3] = {pu, pv, pw};
        double* val[3] = {h->u, h->v, h->w};
        char* mk[3] = {h->um, h->vm, h->wm};
        for (int q = 0; q < 3; ++q) {
            const size_t o = zlo * pl[q], c = (zhi - zlo + (q == 2 ? 1 : 0)) * pl[q];
            JC_TRY(jc_zero(ctx, val[q] + o, c)); JC_TRY(jc_zero(ctx, h->wsum[q] + o, c)); JC_TRY(jc_zero(ctx, mk[q] + o, c));
        }
    }
    if (h->nglobal > 0) {
        if (h->nsorted > 0) {
            P2GOut out;
            out.val[0] = h->u; out.val[1] = h->v; out.val[2] = h->w;
            for (int q = 0; q < 3; ++q) out.wsum[q] = h->wsum[q];
            out.marker[0] = h->um; out.marker[1] = h->vm; out.marker[2] = h->wm;
            // tiles are aligned to the GLOBAL lattice (not to the slab), so that every rank forms the same tile sums.
            static const bool window_kernel = getenv("JETCU_P2G_KERNEL") && !strcmp(getenv("JETCU_P2G_KERNEL"), "window");
            if (!window_kernel) {
                // staggered-lattice kernel: tiles of 8 x 8 x 4 staggered cells, 256 threads
                constexpr int TZ = 4;
                const int ntx = (g.nx + P2G_TX - 1) / P2G_TX, nty = (g.ny + P2G_TY - 1) / P2G_TY;
                const int tzlo = max(kb - 1, 0) / TZ, tzhi = (ke - 1) / TZ;  // staggered planes [kb - 1, ke - 1]
                static const int minb = getenv("JETCU_P2G_MINB") ? atoi(getenv("JETCU_P2G_MINB")) : 0;
                void (*kern)(GridD, int, int, int, int, int, const uint32_t*, const uint32_t*, AttrPtrs, P2GOut, int*) =
                    apic ? (minb == 3 ? k_p2g_stag<true, TZ, 3> : k_p2g_stag<true, TZ, 2>)
                         : (minb == 4 ? k_p2g_stag<false, TZ, 4> : (minb == 2 ? k_p2g_stag<false, TZ, 2> : k_p2g_stag<false, TZ, 3>));
                for (int colour = 0; colour < 8; ++colour) {  // fixed colour order = the order in which a rim face receives its tiles
                    const int par_x = colour & 1, par_y = (colour >> 1) & 1, par_z = colour >> 2;
                    const int cx = (ntx - par_x + 1) / 2, cy = (nty - par_y + 1) / 2;
                    const int tz_first = tzlo + ((tzlo & 1) != par_z ? 1 : 0);
                    const int cz = tz_first > tzhi ? 0 : (tzhi - tz_first) / 2 + 1;
                    if (cx <= 0 || cy <= 0 || cz <= 0) continue;
                    JC_LAUNCH(ctx, kern, dim3(cx, cy, cz), P2G_TX * P2G_TY * TZ, 0, g, par_x, par_y, tz_first, kb, ke, h->cell_start,
                              h->sub_start, a, out, h->p2g_err);
                }
            } else {
            // 18-slot window kernel (round 2's first form, kept for A/B runs).
            // Tile height: 4 cell planes (256 threads) held to 128 registers, so that two CTAs share an SM and one tile's
            // exchange / gather / read-modify-write phases overlap the other's particle loop, with software-pipelined particle
            // loads (GPU call 8, 256^3 FLIP: 12.24 ms for the 8-plane tile without either -> 11.16 ms).  APIC carries 3 more
            // attributes per component through the cell loop and spills at 128 registers: one 256-thread CTA per SM at 227
            // registers, pipelined loads (23.8 -> 19.5 ms).  JETCU_P2G_TZ=8 / JETCU_P2G_PF=0 / JETCU_P2G_MINB select the others.
            static const int tz_env = getenv("JETCU_P2G_TZ") ? atoi(getenv("JETCU_P2G_TZ")) : 0;
            static const int minb_env = getenv("JETCU_P2G_MINB") ? atoi(getenv("JETCU_P2G_MINB")) : 0;
            static const int pf_env = getenv("JETCU_P2G_PF") ? atoi(getenv("JETCU_P2G_PF")) : P2G_PF_DEFAULT;
            const int TZ = (apic || tz_env != P2G_TZ) ? P2G_TZ_APIC : P2G_TZ;   // default since GPU call 8: the 4-plane tile
            const int ntx = (g.nx + P2G_TX - 1) / P2G_TX, nty = (g.ny + P2G_TY - 1) / P2G_TY;
            const int tzlo = kb / TZ, tzhi = (ke - 1) / TZ;  // inclusive
            const size_t smem = (size_t)2 * P2G_SLOTS * P2G_TX * P2G_TY * TZ * sizeof(double);
            const int threads = P2G_TX * P2G_TY * TZ;
            // per device, cheap: set on every call (the attribute belongs to the current device's copy of the function)
            void (*kern)(GridD, int, int, int, int, int, const uint32_t*, AttrPtrs, P2GOut, int*) =
                apic ? (minb_env == 2 ? k_p2g_tile<true, P2G_TZ_APIC, 2, true>
                                      : (pf_env ? k_p2g_tile<true, P2G_TZ_APIC, 0, true> : k_p2g_tile<true, P2G_TZ_APIC, 0, false>))
                     : (TZ == P2G_TZ ? (pf_env ? k_p2g_tile<false, P2G_TZ, 0, true> : k_p2g_tile<false, P2G_TZ, 0, false>)
                                     : (pf_env ? k_p2g_tile<false, P2G_TZ_APIC, 2, true> : k_p2g_tile<false, P2G_TZ_APIC, 2, false>));
            JC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (int colour = 0; colour < 8; ++colour) {  // fixed colour order = the order in which a rim face receives its tiles
                const int par_x = colour & 1, par_y = (colour >> 1) & 1, par_z = colour >> 2;
                const int cx = (ntx - par_x + 1) / 2, cy = (nty - par_y + 1) / 2;
                const int tz_first = tzlo + ((tzlo & 1) != par_z ? 1 : 0);
                const int cz = tz_first > tzhi ? 0 : (tzhi - tz_first) / 2 + 1;
                if (cx <= 0 || cy <= 0 || cz <= 0) continue;
                dim3 grid(cx, cy, cz);
                JC_LAUNCH(ctx, kern, grid, threads, smem, g, par_x, par_y, tz_first, kb, ke, h->cell_start, a, out, h->p2g_err);
            }
            }
        }
        JC_LAUNCH(ctx, k_p2g_normalise, jc_blocks(nk * pu, 256), 256, 0, h->u + k0 * pu, h->wsum[0] + k0 * pu,
                  (flip && !slab) ? h->ud + k0 * pu : nullptr, nk * pu);
        JC_LAUNCH(ctx, k_p2g_normalise, jc_blocks(nk * pv, 256), 256, 0, h->v + k0 * pv, h->wsum[1] + k0 * pv,
                  (flip && !slab) ? h->vd + k0 * pv : nullptr, nk * pv);
        JC_LAUNCH(ctx, k_p2g_normalise, jc_blocks(nkw * pw, 256), 256, 0, h->w + k0 * pw, h->wsum[2] + k0 * pw,
                  (flip && !slab) ? h->wd + k0 * pw : nullptr, nkw * pw);
    } else if (h->P.mode == JETCU_FLIP) {
        JC_TRY(jc_zero(ctx, h->ud, h->nu)); JC_TRY(jc_zero(ctx, h->vd, h->nv)); JC_TRY(jc_zero(ctx, h->wd, h->nw));
    }
    if (slab) {  // the grid stages that follow (extrapolation, boundary conditions, system build) run replicated
        JC_TRY(hyb_gather_planes(h, h->u, pu * 8, false)); JC_TRY(hyb_gather_planes(h, h->v, pv * 8, false));
        JC_TRY(hyb_gather_planes(h, h->w, pw * 8, true));
        JC_TRY(hyb_gather_planes(h, h->um, pu, false)); JC_TRY(hyb_gather_planes(h, h->vm, pv, false));
        JC_TRY(hyb_gather_planes(h, h->wm, pw, true));
        if (flip) {   // FlipSolver3's float snapshot of the transferred velocity (flip_solver3.cpp:39-47): float(u) of the
                      // gathered field, instead of all-gathering a second (float) copy of every plane
            JC_LAUNCH(ctx, k_snapshot_f32, jc_blocks(h->nu, 256), 256, 0, h->u, h->ud, h->nu);
            JC_LAUNCH(ctx, k_snapshot_f32, jc_blocks(h->nv, 256), 256, 0, h->v, h->vd, h->nv);
            JC_LAUNCH(ctx, k_snapshot_f32, jc_blocks(h->nw, 256), 256, 0, h->w, h->wd, h->nw);
        }
    }
    jc_prof_end(ctx, slot, 6, h->np);
    return JETCU_OK;
}

int jetcu_hybrid_build_signed_distance_field(jetcu_hybrid* h) {
    HYB_ENTER(h);
    JC_TRY(hyb_bin(h));
    GridD g = gridd(h);
    double maxH = fmax(fmax(g.hx, g.hy), g.hz);
    double radius = 1.2 * maxH / sqrt(2.0);
    double band = 2.0 * radius;
    // a particle in cell c+d lies at least (|d|-0.5)h from the centre of cell c along that axis
    int rx = (int)floor(band / g.hx + 0.5 + 1e-9), ry = (int)floor(band / g.hy + 0.5 + 1e-9), rz = (int)floor(band / g.hz + 0.5 + 1e-9);
    if (h->world > 1 && rz > SLAB_GHOST)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "particle SDF band reaches %d planes; slab mode carries %d ghost planes", rz,
                          SLAB_GHOST);
    dim3 grid((g.nx + 31) / 32, (g.ny + 3) / 4, h->z1 - h->z0);  // this rank's planes (all of them on a single GPU)
    int c = h->cur;
    JC_LAUNCH(ctx, k_build_sdf, grid, dim3(32, 4), 0, g, h->cell_start, h->attr[c][0], h->attr[c][1], h->attr[c][2], radius,
              band, rx, ry, rz, h->z0, h->sdf);
    JC_TRY(hyb_gather_planes(h, h->sdf, (size_t)h->nx * h->ny * 8, false));
    if (h->col.kind != 0) {  // GridFluidSolver3::extrapolateIntoCollider(ScalarGrid3*), grid_fluid_solver3.cpp:355-369
        JC_LAUNCH(ctx, k_outside_collider_marker, face_grid(h->nx, h->ny, h->nz), dim3(64, 4), 0, g, h->col_sdf, h->cm_u);
        JC_TRY(hyb_extrapolate_one(h, h->sdf, h->cm_u, h->nx, h->ny, h->nz, (unsigned)ceil(h->P.max_cfl)));
    }
    // extrapolateIntoCollider (src/jet/grid_fluid_solver3.cpp:355-369) is the identity without a collider
    return JETCU_OK;
}

int jetcu_hybrid_extrapolate_velocity_to_air(jetcu_hybrid* h) {
    HYB_ENTER(h);
    unsigned depth = (unsigned)ceil(h->P.max_cfl);  // src/jet/pic_solver3.cpp:284-294
    JC_TRY(hyb_extrapolate_one(h, h->u, h->um, h->nx + 1, h->ny, h->nz, depth));
    JC_TRY(hyb_extrapolate_one(h, h->v, h->vm, h->nx, h->ny + 1, h->nz, depth));
    return hyb_extrapolate_one(h, h->w, h->wm, h->nx, h->ny, h->nz + 1, depth);
}

int jetcu_hybrid_begin_advance_time_step(jetcu_hybrid* h, double dt) {
    HYB_ENTER(h);
    (void)dt;
    // GridFluidSolver3::beginAdvanceTimeStep (:437-462) then PicSolver3::onBeginAdvanceTimeStep (pic_solver3.cpp:63-93)
    JC_TRY(hyb_zero_walls(h));
    JC_TRY(jetcu_hybrid_transfer_particles_to_grids(h));
    JC_TRY(jetcu_hybrid_build_signed_distance_field(h));
    JC_TRY(jetcu_hybrid_extrapolate_velocity_to_air(h));
    return hyb_zero_walls(h);
}

int jetcu_hybrid_compute_external_forces(jetcu_hybrid* h, double dt) {
    HYB_ENTER(h);
    const double* g = h->P.gravity;
    if (g[0] * g[0] + g[1] * g[1] + g[2] * g[2] > DBL_EPSILON) {
        if (fabs(g[0]) > DBL_EPSILON) JC_LAUNCH(ctx, k_add_scalar, jc_blocks(h->nu, 256), 256, 0, h->u, dt * g[0], h->nu);
        if (fabs(g[1]) > DBL_EPSILON) JC_LAUNCH(ctx, k_add_scalar, jc_blocks(h->nv, 256), 256, 0, h->v, dt * g[1], h->nv);
        if (fabs(g[2]) > DBL_EPSILON) JC_LAUNCH(ctx, k_add_scalar, jc_blocks(h->nw, 256), 256, 0, h->w, dt * g[2], h->nw);
        JC_TRY(hyb_zero_walls(h));
    }
    return JETCU_OK;
}

// GridFluidSolver3::setViscosityCoefficient / computeViscosity (src/jet/grid_fluid_solver3.cpp:231-241): the solver's own
// GridBackwardEulerDiffusionSolver3 (default boundary type Neumann, ICCG(100, kEpsilonD)) on a clone of the velocity,
// boundarySdf = colliderSdf(), fluidSdf = the particle SDF; then applyBoundaryCondition.
int jetcu_hybrid_set_viscosity(jetcu_hybrid* h, double coefficient) {
    if (!h) return JETCU_ERR_INVALID_ARG;
    if (!(coefficient >= 0.0)) return jetcu_fail(h->ctx, JETCU_ERR_INVALID_ARG, "viscosity coefficient %g", coefficient);
    h->viscosity = coefficient;
    return JETCU_OK;
}

int jetcu_hybrid_compute_viscosity(jetcu_hybrid* h, double dt, jetcu_mgpcg_result results[3]) {
    HYB_ENTER(h);
    if (results) memset(results, 0, 3 * sizeof(jetcu_mgpcg_result));
    if (!(h->viscosity > DBL_EPSILON)) return JETCU_OK;
    if (h->world > 1)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "computeViscosity (ICCG) is not slab-distributed: one GPU only");
    double *u0 = nullptr, *v0 = nullptr, *w0 = nullptr;
    int rc = JETCU_OK;
#define STEP(e) if (rc == JETCU_OK) rc = (e)
    STEP(jc_alloc(ctx, &u0, h->nu)); STEP(jc_alloc(ctx, &v0, h->nv)); STEP(jc_alloc(ctx, &w0, h->nw));
    STEP(jc_d2d(ctx, u0, h->u, h->nu)); STEP(jc_d2d(ctx, v0, h->v, h->nv)); STEP(jc_d2d(ctx, w0, h->w, h->nw));
    STEP(diffusion_solve_faces(ctx, gridd(h), u0, v0, w0, h->viscosity, dt, 1, h->col.kind != 0 ? h->col_sdf : nullptr, h->sdf, 100,
                               DBL_EPSILON, h->u, h->v, h->w, results));
#undef STEP
    jc_free(ctx, u0); jc_free(ctx, v0); jc_free(ctx, w0);
    if (rc != JETCU_OK) return rc;
    return jetcu_hybrid_apply_boundary_condition(h);
}

int jetcu_hybrid_compute_pressure(jetcu_hybrid* h, double dt, jetcu_mgpcg_result* result) {
    HYB_ENTER(h);
    (void)dt;
    // GridFluidSolver3::computePressure (:243-253): input is a clone of the velocity; every face is read and written
    // by the same thread of k_apply_gradient, so solving in place is equivalent.
    const bool col = h->col.kind != 0;
    JC_TRY(pressure_solve_device(ctx, &h->ps, h->u, h->v, h->w, col ? h->col_sdf : nullptr, col ? h->bnd_u : nullptr,
                                 col ? h->bnd_v : nullptr, col ? h->bnd_w : nullptr, h->sdf, &h->P.mgpcg, h->u, h->v, h->w,
                                 &h->last));
    if (result) *result = h->last;
    return hyb_zero_walls(h);
}

int jetcu_hybrid_transfer_grids_to_particles(jetcu_hybrid* h) {
    HYB_ENTER(h);
    GridD g = gridd(h);
    size_t np = h->np;
    AttrPtrs a = owned_ptrs(h);
    if (h->P.mode == JETCU_FLIP) {
        JC_LAUNCH(ctx, k_flip_delta, jc_blocks(h->nu, 256), 256, 0, h->u, h->ud, h->nu);
        JC_LAUNCH(ctx, k_flip_delta, jc_blocks(h->nv, 256), 256, 0, h->v, h->vd, h->nv);
        JC_LAUNCH(ctx, k_flip_delta, jc_blocks(h->nw, 256), 256, 0, h->w, h->wd, h->nw);
        if (np) JC_LAUNCH(ctx, k_g2p_flip, jc_blocks(np, 128), 128, 0, g, h->u, h->v, h->w, h->ud, h->vd, h->wd,
                          h->P.pic_blending_factor, a, np);
    } else if (h->P.mode == JETCU_APIC) {
        if (np) HYB_SAMPLER_LAUNCH(k_g2p_apic, jc_blocks(np, 128), 128, g, h->u, h->v, h->w, a, np);
    } else {
        if (np) HYB_SAMPLER_LAUNCH(k_g2p_pic, jc_blocks(np, 128), 128, g, h->u, h->v, h->w, a, np);
    }
    return JETCU_OK;
}

// pair arrays of the current velocity field; flip_delta: also turn FLIP's snapshot into the delta field (one pass per field)
static int hyb_pack_pairs(jetcu_hybrid* h, bool flip_delta) {
    jetcu_ctx* ctx = h->ctx;
    JC_LAUNCH(ctx, k_pack_pairs, jc_blocks(h->nu, 256), 256, 0, h->u, h->pair[0], flip_delta ? h->ud : nullptr, h->nx + 1, h->nu);
    JC_LAUNCH(ctx, k_pack_pairs, jc_blocks(h->nv, 256), 256, 0, h->v, h->pair[1], flip_delta ? h->vd : nullptr, h->nx, h->nv);
    JC_LAUNCH(ctx, k_pack_pairs, jc_blocks(h->nw, 256), 256, 0, h->w, h->pair[2], flip_delta ? h->wd : nullptr, h->nx, h->nw);
    return JETCU_OK;
}

// whether the advection kernel of this step also does the next binning's counting pass: one GPU (slab mode routes and
// exchanges particles first), no collider (resolveCollision moves particles after the advection kernel)
static bool hyb_count_in_move(const jetcu_hybrid* h) {
    return h->world == 1 && h->col.kind == 0 && h->np > 0 && !getenv("JETCU_NO_FUSED_COUNT");
}

static int hyb_before_move(jetcu_hybrid* h, bool count) {
    if (!count) return JETCU_OK;
    return jc_zero(h->ctx, h->sub_start, 8 * h->nc + 2);
}

static int hyb_after_move(jetcu_hybrid* h, bool counted) {
    jetcu_ctx* ctx = h->ctx;
    if (h->np && h->col.kind != 0)
        JC_LAUNCH(ctx, k_resolve_collision, jc_blocks(h->np, 128), 128, 0, h->col, owned_ptrs(h), (size_t)h->np);
    h->nsorted = h->np;  // the ghost copies are stale from here on
    h->binned = false;
    h->counted = counted;
    return JETCU_OK;
}

int jetcu_hybrid_move_particles(jetcu_hybrid* h, double dt) {
    HYB_ENTER(h);
    GridD g = gridd(h);
    unsigned nsub = (unsigned)fmax(h->P.max_cfl, 1.0);
    double sub_dt = dt / nsub;
    const bool count = hyb_count_in_move(h);
    JC_TRY(hyb_before_move(h, count));
    if (h->np) {
        JC_TRY(hyb_pack_pairs(h, false));
        HYB_MOVE_LAUNCH(k_move, 2, jc_blocks(h->np, 128), 128, g, h->pair[0], h->pair[1], h->pair[2], sub_dt, nsub,
                           h->P.closed_domain_boundary_flag, owned_ptrs(h), (size_t)h->np, h->cell_of, h->rank,
                           count ? h->sub_start : nullptr);
    }
    return hyb_after_move(h, count);
}

// transferFromGridsToParticles + moveParticles of one sub-step; PIC and FLIP (blending factor 0) in one fused kernel
static int hyb_g2p_move(jetcu_hybrid* h, double dt) {
    jetcu_ctx* ctx = h->ctx;
    const bool fused = h->P.mode == JETCU_PIC || (h->P.mode == JETCU_FLIP && !(h->P.pic_blending_factor > 0.0));
    if (!fused || getenv("JETCU_UNFUSED_ADVECTION")) {
        JC_TRY(jetcu_hybrid_transfer_grids_to_particles(h));
        return jetcu_hybrid_move_particles(h, dt);
    }
    GridD g = gridd(h);
    const unsigned nsub = (unsigned)fmax(h->P.max_cfl, 1.0);
    const double sub_dt = dt / nsub;
    const bool flip = h->P.mode == JETCU_FLIP;
    JC_TRY(hyb_pack_pairs(h, flip));
    const bool count = hyb_count_in_move(h);
    JC_TRY(hyb_before_move(h, count));
    if (h->np) {
        if (flip)
            HYB_MOVE_LAUNCH(k_flip_move, 5, jc_blocks(h->np, 128), 128, g, h->pair[0], h->pair[1], h->pair[2], h->ud, h->vd, h->wd,
                               sub_dt, nsub, h->P.closed_domain_boundary_flag, owned_ptrs(h), (size_t)h->np, h->cell_of, h->rank,
                               count ? h->sub_start : nullptr);
        else
            HYB_MOVE_LAUNCH(k_pic_move, 2, jc_blocks(h->np, 128), 128, g, h->pair[0], h->pair[1], h->pair[2], sub_dt, nsub,
                               h->P.closed_domain_boundary_flag, owned_ptrs(h), (size_t)h->np, h->cell_of, h->rank,
                               count ? h->sub_start : nullptr);
    }
    return hyb_after_move(h, count);
}

int jetcu_hybrid_compute_advection(jetcu_hybrid* h, double dt) {
    HYB_ENTER(h);
    // PicSolver3::computeAdvection, src/jet/pic_solver3.cpp:95-112
    JC_TRY(jetcu_hybrid_extrapolate_velocity_to_air(h));
    JC_TRY(hyb_zero_walls(h));
    return hyb_g2p_move(h, dt);
}

int jetcu_hybrid_step(jetcu_hybrid* h, double dt, jetcu_step_stats* st) {
    HYB_ENTER(h);
    cudaStream_t s = ctx->stream;
#define MARK(i) cudaEventRecord(h->ev[i], s)
    MARK(0);
    JC_TRY(hyb_zero_walls(h));
    JC_TRY(hyb_bin(h));
    MARK(1);
    JC_TRY(jetcu_hybrid_transfer_particles_to_grids(h));
    MARK(2);
    JC_TRY(jetcu_hybrid_build_signed_distance_field(h));
    MARK(3);
    JC_TRY(jetcu_hybrid_extrapolate_velocity_to_air(h));
    MARK(4);
    JC_TRY(hyb_zero_walls(h));
    JC_TRY(jetcu_hybrid_compute_external_forces(h, dt));
    JC_TRY(jetcu_hybrid_compute_viscosity(h, dt, nullptr));   // no-op unless set_viscosity(> kEpsilonD); counted under ms_bc
    MARK(5);
    JC_TRY(jetcu_hybrid_compute_pressure(h, dt, nullptr));
    MARK(6);
    JC_TRY(jetcu_hybrid_extrapolate_velocity_to_air(h));
    JC_TRY(hyb_zero_walls(h));
    MARK(7);
    MARK(8);   // G2P and advection run fused (hyb_g2p_move): ms_g2p is reported as 0, ms_move covers both
    JC_TRY(hyb_g2p_move(h, dt));
    MARK(9);
#undef MARK
    int p2g_err = 0;
    JC_TRY(jc_d2h(ctx, &p2g_err, h->p2g_err, 1));
    JC_TRY(jc_sync(ctx));
    if (p2g_err) {
        jc_zero(ctx, h->p2g_err, 1);
        return jetcu_fail(ctx, JETCU_ERR_CUDA,
                          "P2G: a particle lies outside the face window of the cell it was binned into");
    }
    if (st) {
        memset(st, 0, sizeof(*st));
        auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]); return (double)ms; };
        st->pcg_iterations = h->last.iterations;
        st->pcg_last_residual = h->last.last_residual;
        st->pcg_true_residual_l2 = h->last.true_residual_l2;
        st->pcg_b_l2 = h->last.b_l2;
        st->ms_total = el(0, 9);
        st->ms_bin = el(0, 1);
        st->ms_p2g = el(1, 2);
        st->ms_sdf = el(2, 3);
        st->ms_extrapolate = el(3, 4) + el(6, 7);
        st->ms_bc = el(4, 5);
        st->ms_pressure = el(5, 6);
        st->ms_g2p = el(7, 8);
        st->ms_move = el(8, 9);
        st->ms_mgpcg_solve = h->last.solve_ms;
    }
    return JETCU_OK;
}

int jetcu_hybrid_advance_frame(jetcu_hybrid* h, double frame_dt, uint32_t* n_sub_steps, jetcu_step_stats* last) {
    HYB_ENTER(h);
    // PhysicsAnimation::advanceTimeStep, adaptive branch (src/jet/physics_animation.cpp:102-121)
    double remaining = frame_dt;
    uint32_t count = 0;
    while (remaining > DBL_EPSILON) {
        uint32_t ns = 1;
        JC_TRY(jetcu_hybrid_num_sub_steps(h, remaining, &ns));
        double actual = remaining / (double)ns;
        JC_TRY(jetcu_hybrid_step(h, actual, last));
        remaining -= actual;
        ++count;
    }
    if (n_sub_steps) *n_sub_steps = count;
    return JETCU_OK;
}

}  // extern "C"
