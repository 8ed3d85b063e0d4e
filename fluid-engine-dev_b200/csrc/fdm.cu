// fdm.cu — sm_100a kernels for Jet's FDM linear-system chain and the MGPCG driver.
//
// Every kernel replays the reference's floating-point expression order (cited per kernel) and the TU is built
// with -fmad=false, so mvm / residual / red-black SOR / restrict / correct are BIT-IDENTICAL to the CPU path.
// Only the dot product differs (tree instead of a 16M-long serial chain) — deterministic, fixed shape.
//
// Device layout: matrix as 4 SoA planes (MatSoA), vectors as plain nx*ny*nz doubles, x fastest.
#include "fdm.cuh"
#include "fdm_ops.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>


// ------------------------------------------------------------------------------------------------ AoS <-> SoA
__global__ void k_aos_to_soa(const jetcu_matrix_row* __restrict__ aos, double* __restrict__ c, double* __restrict__ r,
                             double* __restrict__ u, double* __restrict__ f, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2* p = reinterpret_cast<const double2*>(aos + i);
    double2 a = p[0], b = p[1];
    c[i] = a.x; r[i] = a.y; u[i] = b.x; f[i] = b.y;
}

template <bool P>
__global__ void k_mat_to_aos(Mat A, jetcu_matrix_row* __restrict__ aos, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2* p = reinterpret_cast<double2*>(aos + i);
    p[0] = make_double2(A.center[i], m_right<P>(A, i));
    p[1] = make_double2(m_up<P>(A, i), m_front<P>(A, i));
}

int fdm_alloc_mat(jetcu_ctx* ctx, MatSoA* A, size_t n) {
    double* base;
    JC_TRY(jc_alloc(ctx, &base, 4 * n));
    A->center = base; A->right = base + n; A->up = base + 2 * n; A->front = base + 3 * n;
    return JETCU_OK;
}

void fdm_free_mat(jetcu_ctx* ctx, MatSoA* A) {
    jc_free(ctx, A->center);
    A->right = A->up = A->front = nullptr;
}

int fdm_aos_to_soa(jetcu_ctx* ctx, const jetcu_matrix_row* d_aos, MatSoA A, size_t n) {
    JC_LAUNCH(ctx, k_aos_to_soa, jc_blocks(n, 256), 256, 0, d_aos, A.center, A.right, A.up, A.front, n);
    return JETCU_OK;
}

int fdm_mat_to_aos(jetcu_ctx* ctx, Mat A, jetcu_matrix_row* d_aos, size_t n) {
    if (A.packed) JC_LAUNCH(ctx, k_mat_to_aos<true>, jc_blocks(n, 256), 256, 0, A, d_aos, n);
    else JC_LAUNCH(ctx, k_mat_to_aos<false>, jc_blocks(n, 256), 256, 0, A, d_aos, n);
    return JETCU_OK;
}

// ------------------------------------------------------------------------------------------------ stencils
#define BX 64
#define BY 4

// Kernel bodies (one cell each) live in fdm_ops.cuh as __host__ __device__ functions; tests/hostcheck runs them on the CPU.
template <bool P>
__global__ void __launch_bounds__(BX* BY) k_mvm(Mat A, const double* __restrict__ v, double* __restrict__ out,
                                                int nx, int ny, int nz, int k0) {
    int i = blockIdx.x * BX + threadIdx.x, j = blockIdx.y * BY + threadIdx.y, k = k0 + blockIdx.z;
    if (i >= nx || j >= ny) return;
    mvm_cell<P>(A, v, out, i, j, k, nx, ny, nz);
}

template <bool P>
__global__ void __launch_bounds__(BX* BY) k_residual(Mat A, const double* __restrict__ x,
                                                     const double* __restrict__ b, double* __restrict__ out, int nx,
                                                     int ny, int nz, int k0) {
    int i = blockIdx.x * BX + threadIdx.x, j = blockIdx.y * BY + threadIdx.y, k = k0 + blockIdx.z;
    if (i >= nx || j >= ny) return;
    residual_cell<P>(A, x, b, out, i, j, k, nx, ny, nz);
}

// ---- FdmGaussSeidelSolver3::relaxRedBlack, src/jet/fdm_gauss_seidel_solver3.cpp:150-221 -----------------------
// (a) Small grids: ONE CTA sweeps the whole level in place, `iterations` times, with __syncthreads() between the
//     colour passes — the coarse MG levels are pure launch latency otherwise.
#define SMALL_THREADS 1024
#define SMALL_MAX_CELLS 8192
//     SMEM: x lives in shared memory for the duration (levels of at most SMALL_SMEM_CELLS cells = 48 KB): a colour pass then
//     costs a shared-memory round trip instead of a global one per __syncthreads (the same arithmetic on the same values).
#define SMALL_SMEM_CELLS 6144
template <bool P, bool SMEM>
__global__ void __launch_bounds__(SMALL_THREADS) k_relax_small(Mat A, const double* __restrict__ b, double* x, double w,
                                                               int nx, int ny, int nz, unsigned iterations) {
    __shared__ double xs[SMEM ? SMALL_SMEM_CELLS : 1];
    const int hx = (nx + 1) / 2;
    const int pairs = hx * ny * nz;
    double* xw = x;
    if (SMEM) {
        for (int c = threadIdx.x; c < nx * ny * nz; c += SMALL_THREADS) xs[c] = x[c];
        __syncthreads();
        xw = xs;
    }
    for (unsigned it = 0; it < 2 * iterations; ++it) {
        const int colour = it & 1;
        for (int t = threadIdx.x; t < pairs; t += SMALL_THREADS) {
            int ip = t % hx, j = (t / hx) % ny, k = t / (hx * ny);
            int i = 2 * ip + ((j + k + colour) & 1);
            if (i >= nx) continue;
            relax_cell<P>(A, b, xw, w, i, j, k, nx, ny, nz);
        }
        __syncthreads();
    }
    if (SMEM)
        for (int c = threadIdx.x; c < nx * ny * nz; c += SMALL_THREADS) x[c] = xs[c];
}

// (b) Any grid, any matrix form: one launch per colour, one thread per cell of that colour (x-pairs).  With 8-byte
//     off-diagonal planes (host-supplied matrices) a sweep moves ~110 B/cell, because every 32-byte sector of every
//     plane is fetched in both passes with half of it unused (ncu: profiles/r01_baseline_relax_colour_full.txt).
template <bool P>
__global__ void __launch_bounds__(BX* BY) k_relax_colour(Mat A, const double* __restrict__ b, double* __restrict__ x,
                                                         double w, int colour, int nx, int ny, int nz, int k0) {
    jc_pdl_prologue();
    int j = blockIdx.y * BY + threadIdx.y, k = k0 + blockIdx.z;
    int i = 2 * (blockIdx.x * BX + threadIdx.x) + ((j + k + colour) & 1);
    if (i >= nx || j >= ny) return;
    relax_cell<P>(A, b, x, w, i, j, k, nx, ny, nz);
}

// (b') Packed matrix with face codes: the same two colour passes, but the three 8-byte off-diagonal planes are
//     replaced by one code byte per cell (2 bits per +face: 0 closed, 1 fully open => -invH^2 with no FP64 work,
//     2 fractional => weight fetched on demand), so a pass streams center 8 + b 8 + x 8 r + ~8 w + code 1 = 33 B/cell.
__global__ void __launch_bounds__(BX* BY) k_relax_colour_code(Mat A, const double* __restrict__ b, double* __restrict__ x,
                                                              double w, int colour, int nx, int ny, int nz, int k0) {
    jc_pdl_prologue();
    const int j = blockIdx.y * BY + threadIdx.y, k = k0 + blockIdx.z;
    const int i = 2 * (blockIdx.x * BX + threadIdx.x) + ((j + k + colour) & 1);
    if (i >= nx || j >= ny) return;
    relax_cell_code(A, b, x, w, i, j, k, nx, ny, nz);
}

// A 3-D CTA tile (TX x-pairs x TY rows x TZ planes, so that the z-neighbours of the inner planes hit in the L1) was measured in
// round 2 (GPU call 17, profiles/r02_sor_wavefront.md): eight shapes from 64x4x2 to 32x4x8, none faster than the plane tile
// (25.28-35.9 ms per solve against 25.38) — the pass is bound by DRAM, not by L2 -> L1 traffic.  Removed again.

static inline dim3 grid3z(Dim3i d, ZRange z) { return dim3((d.nx + BX - 1) / BX, (d.ny + BY - 1) / BY, z.k1 - z.k0); }
static inline ZRange zfull(Dim3i d) { return ZRange{0, d.nz}; }

template <bool P>
static int launch_mvm(jetcu_ctx* ctx, Dim3i d, ZRange z, Mat A, const double* v, double* out) {
    JC_LAUNCH(ctx, k_mvm<P>, grid3z(d, z), dim3(BX, BY), 0, A, v, out, d.nx, d.ny, d.nz, z.k0);
    return JETCU_OK;
}

int fdm_mvm(jetcu_ctx* ctx, Dim3i d, ZRange z, Mat A, const double* v, double* out) {
    if (z.k1 <= z.k0) return JETCU_OK;
    return A.packed ? launch_mvm<true>(ctx, d, z, A, v, out) : launch_mvm<false>(ctx, d, z, A, v, out);
}
int fdm_mvm(jetcu_ctx* ctx, Dim3i d, Mat A, const double* v, double* out) { return fdm_mvm(ctx, d, zfull(d), A, v, out); }

template <bool P>
static int launch_residual(jetcu_ctx* ctx, Dim3i d, ZRange z, Mat A, const double* x, const double* b, double* out) {
    JC_LAUNCH(ctx, k_residual<P>, grid3z(d, z), dim3(BX, BY), 0, A, x, b, out, d.nx, d.ny, d.nz, z.k0);
    return JETCU_OK;
}

int fdm_residual(jetcu_ctx* ctx, Dim3i d, ZRange z, Mat A, const double* x, const double* b, double* out) {
    if (z.k1 <= z.k0) return JETCU_OK;
    return A.packed ? launch_residual<true>(ctx, d, z, A, x, b, out) : launch_residual<false>(ctx, d, z, A, x, b, out);
}
int fdm_residual(jetcu_ctx* ctx, Dim3i d, Mat A, const double* x, const double* b, double* out) {
    return fdm_residual(ctx, d, zfull(d), A, x, b, out);
}

// One colour pass over the planes of `z` (colour 0 = red = (i+j+k) even with GLOBAL k).
// `after_pass` = the previous operation of the stream is another colour pass (launch_relax keeps track): the launch may then
// be a programmatic dependent launch (common.cuh), which hides the launch gap between the ~50 dependent passes of a level visit.
template <bool P>
static int launch_colour(jetcu_ctx* ctx, Dim3i d, ZRange z, Mat A, const double* b, double sor, int colour, double* x,
                         bool after_pass = false) {
    dim3 grid(((d.nx + 1) / 2 + BX - 1) / BX, (d.ny + BY - 1) / BY, z.k1 - z.k0);
    if (P && A.code != nullptr)
        JC_LAUNCH_PDL(ctx, after_pass, k_relax_colour_code, grid, dim3(BX, BY), 0, A, b, x, sor, colour, d.nx, d.ny, d.nz, z.k0);
    else
        JC_LAUNCH_PDL(ctx, after_pass, k_relax_colour<P>, grid, dim3(BX, BY), 0, A, b, x, sor, colour, d.nx, d.ny, d.nz, z.k0);
    return JETCU_OK;
}

// `halo` = slab-distributed level.  Communication-avoiding sweeps: instead of refreshing one plane before every colour
// pass (2 messages per neighbour and sweep), a block of k sweeps starts with ONE exchange of 2k planes and then
// recomputes the neighbour's cells inside that halo redundantly, on a window that shrinks by one plane per colour pass
// (red of sweep s on slab +- (2(k-s)-1) planes, black on slab +- (2(k-s)-2)); the last black pass covers exactly the
// slab.  Every cell still sees the same operands in the same order, so the result stays bit-identical to the
// undivided sweep; the redundant work is ~k/T of a pass, k is chosen as T/8 (at most 8, at least 1).  The caller must
// have made b valid on fdm_relax_rhs_halo(T) planes around the slab (b is read-only here: one exchange per V-cycle level).
// sweeps per exchange on a slab of T planes
// sweeps per exchange on a slab of T planes with `plane_cells` cells per plane.  One exchange costs a latency L (~25 us for
// a dependent NCCL send/recv pair, profiles/r02_multi) whatever its depth — the BYTES per sweep are 2 planes per neighbour
// either way — while the redundant halo work of a block of k sweeps is ~4 k planes per sweep: k* = sqrt(L / (4 t_plane)),
// t_plane = plane_cells x 23.6 B / HBM rate (ncu: 23.6 B per cell and colour pass).  256^2 planes: 5; 512^2: 3; 128^2: 8.
// At most 8 and at most T / 4 (the halo of 2 k planes must come out of the neighbour's own slab).
// JETCU_RELAX_BLOCK_DIV = d forces the round-1 rule k = T / d.
static inline int relax_block_sweeps(int T, size_t plane_cells) {
    static int div = -1;
    if (div < 0) { const char* e = getenv("JETCU_RELAX_BLOCK_DIV"); div = e ? atoi(e) : 0; if (div < 0) div = 0; }
    int k;
    if (div >= 2) {
        k = T / div;
    } else {
        const double t_plane = (double)plane_cells * 23.6 / 6.5e12;
        k = (int)(sqrt(25e-6 / (4.0 * t_plane)) + 0.5);
        if (k > T / 4) k = T / 4;
    }
    return k < 1 ? 1 : (k > 8 ? 8 : k);
}
// planes of b (which the sweeps only read) that the redundant halo cells need next to a slab of T planes
int fdm_relax_rhs_halo(int T, size_t plane_cells) { return 2 * relax_block_sweeps(T, plane_cells) - 1; }

// planes per chunk for a level with `plane_cells` cells per plane: JETCU_RELAX_CHUNK (planes) if set, else the number of
// planes that makes a launch ~3.5 M cells (JETCU_RELAX_CHUNK_CELLS) — 48-64 planes at 256^2, 13 at 512^2, 6 at 768^2: the
// measured optimum at 256^3 (profiles/r02_sor_wavefront.md) carried over by launch size.  0 = whole-level passes.
static int relax_chunk_planes(size_t plane_cells) {
    static long long planes = -2, cells = -1;
    if (planes == -2) {
        const char* e = getenv("JETCU_RELAX_CHUNK");
        planes = e ? atoll(e) : -1;
        const char* c = getenv("JETCU_RELAX_CHUNK_CELLS");
        cells = c ? atoll(c) : 3500000ll;
    }
    if (planes >= 0) return (int)planes;
    if (cells <= 0) return 0;
    const long long p = cells / (long long)(plane_cells ? plane_cells : 1);
    return p < 1 ? 1 : (p > (1 << 20) ? (1 << 20) : (int)p);
}

// one sweep: red on the planes `rw`, black on the planes `bw` (bw inside rw; rw reaches one plane beyond bw wherever the
// level continues), as whole passes (chunk == 0) or chunk-interleaved
template <bool P>
static int launch_sweep(jetcu_ctx* ctx, Dim3i d, ZRange rw, ZRange bw, int chunk, Mat A, const double* b, double sor, double* x,
                        uint64_t* launches, bool after_pass) {
    if (chunk <= 0 || bw.k1 - bw.k0 <= chunk) {
        JC_TRY(launch_colour<P>(ctx, d, rw, A, b, sor, 0, x, after_pass));
        JC_TRY(launch_colour<P>(ctx, d, bw, A, b, sor, 1, x, true));
        *launches += 2;
        return JETCU_OK;
    }
    int red_done = rw.k0;
    for (int c0 = bw.k0; c0 < bw.k1; c0 += chunk) {
        const int c1 = c0 + chunk < bw.k1 ? c0 + chunk : bw.k1;
        const int need = c1 + 1 < rw.k1 ? c1 + 1 : rw.k1;   // black on [c0, c1) reads red up to plane c1
        if (need > red_done) {
            JC_TRY(launch_colour<P>(ctx, d, ZRange{red_done, need}, A, b, sor, 0, x, after_pass));
            red_done = need;
            *launches += 1;
            after_pass = true;
        }
        JC_TRY(launch_colour<P>(ctx, d, ZRange{c0, c1}, A, b, sor, 1, x, after_pass));
        *launches += 1;
        after_pass = true;
    }
    if (red_done < rw.k1) {
        JC_TRY(launch_colour<P>(ctx, d, ZRange{red_done, rw.k1}, A, b, sor, 0, x, after_pass));
        *launches += 1;
    }
    return JETCU_OK;
}

// L2-resident sweeps, several at a time (JETCU_RELAX_SKEW = sweeps per group, default 20, needs a chunk size > 0): the 2S colour
// passes of S consecutive sweeps are issued as a skewed wavefront over z-chunks — for chunk c = 0, 1, ...: for pass
// p = 0 .. 2S-1: colour p & 1 on the planes [c*C - p, (c+1)*C - p).  Pass p reaches plane k only after pass p - 1 covered
// k - 1 .. k + 1 (it is one plane ahead in every chunk), and pass p + 1 has not yet touched k - 1 .. k + 1 (it trails by
// one plane), so every cell update reads exactly the operands of the S full sweeps run one after the other: bit-identical.
// The working set between the first and the last touch of a plane is (C + 2S) planes of x / b / centre / code, which the
// 126 MB L2 holds, so DRAM sees each plane once per GROUP instead of twice per sweep.
static int relax_skew_sweeps() {
    static int c = -1;
    if (c < 0) {
        const char* e = getenv("JETCU_RELAX_SKEW");
        c = e ? atoi(e) : 20;   // measured (profiles/r02_sor_wavefront.md); <= 1 = one sweep at a time
        if (c < 0) c = 0;
    }
    return c;
}

// w0[p] / w1[p] = the plane window of pass p (all equal on an undivided level; on a slab the redundant-halo windows that
// shrink by one plane per pass, see launch_relax).
template <bool P>
static int launch_sweeps_skewed(jetcu_ctx* ctx, Dim3i d, const int* w0, const int* w1, int passes, int chunk, Mat A, const double* b,
                                double sor, double* x, uint64_t* launches, bool after_pass) {
    for (int c0 = w0[0];; c0 += chunk) {
        bool any = false;
        for (int p = 0; p < passes; ++p) {
            if (c0 - p >= w1[p]) continue;      // this pass is done with the level
            any = true;
            const int lo = c0 - p > w0[p] ? c0 - p : w0[p], hi = c0 + chunk - p < w1[p] ? c0 + chunk - p : w1[p];
            if (lo < hi) {
                JC_TRY(launch_colour<P>(ctx, d, ZRange{lo, hi}, A, b, sor, p & 1, x, after_pass));
                *launches += 1;
                after_pass = true;
            }
        }
        if (!any) break;
    }
    return JETCU_OK;
}

template <bool P>
static int launch_relax(jetcu_ctx* ctx, Dim3i d, ZRange z, bool halo, Mat A, const double* b, double sor,
                        unsigned iterations, double** x, double** xalt, uint64_t* launches) {
    (void)xalt;
    if (!halo && d.n() <= SMALL_MAX_CELLS) {
        if (d.n() <= SMALL_SMEM_CELLS) JC_LAUNCH(ctx, (k_relax_small<P, true>), 1, SMALL_THREADS, 0, A, b, *x, sor, d.nx, d.ny, d.nz, iterations);
        else JC_LAUNCH(ctx, (k_relax_small<P, false>), 1, SMALL_THREADS, 0, A, b, *x, sor, d.nx, d.ny, d.nz, iterations);
        *launches = 1;
        return JETCU_OK;
    }
    const size_t plane = (size_t)d.nx * d.ny * sizeof(double);
    const int kmax = relax_block_sweeps(z.k1 - z.k0, (size_t)d.nx * d.ny);
    // chunking only pays where the level does not fit the L2 anyway (>= 64^3-ish: 25 B x 2M cells)
    const int chunk = d.n() >= (size_t)(1 << 21) ? relax_chunk_planes((size_t)d.nx * d.ny) : 0;
    *launches = 0;
    unsigned it = 0;
    const int skew = relax_skew_sweeps();
    int w0[64], w1[64];
    if (!halo && chunk > 0 && skew > 1 && z.k1 - z.k0 > chunk) {   // undivided level: groups of `skew` sweeps as one wavefront
        while (it < iterations) {
            int s2 = (int)((iterations - it < (unsigned)skew) ? iterations - it : (unsigned)skew);
            if (s2 > 32) s2 = 32;
            for (int p = 0; p < 2 * s2; ++p) { w0[p] = z.k0; w1[p] = z.k1; }
            JC_TRY(launch_sweeps_skewed<P>(ctx, d, w0, w1, 2 * s2, chunk, A, b, sor, *x, launches, it > 0));
            it += (unsigned)s2;
        }
        return JETCU_OK;
    }
    while (it < iterations) {
        const int k = halo ? (int)((iterations - it < (unsigned)kmax) ? iterations - it : (unsigned)kmax) : (int)(iterations - it);
        if (halo) JC_TRY(comm_halo_planes(ctx, *x, plane, z.k0, z.k1, d.nz, 2 * k));
        if (halo && chunk > 0 && skew > 1 && z.k1 - z.k0 > chunk && k <= 32) {
            // a slab's block of k sweeps between two exchanges as ONE wavefront: pass p works on the slab +- (2k - 1 - p) planes
            for (int p = 0; p < 2 * k; ++p) {
                const int e = 2 * k - 1 - p;
                w0[p] = z.k0 - e < 0 ? 0 : z.k0 - e; w1[p] = z.k1 + e > d.nz ? d.nz : z.k1 + e;
            }
            JC_TRY(launch_sweeps_skewed<P>(ctx, d, w0, w1, 2 * k, chunk, A, b, sor, *x, launches, false));  // follows the exchange
            it += (unsigned)k;
            continue;
        }
        for (int s2 = 0; s2 < k; ++s2) {
            const int er = halo ? 2 * (k - s2) - 1 : 0, eb = halo ? 2 * (k - s2) - 2 : 0;
            const ZRange rw{z.k0 - er < 0 ? 0 : z.k0 - er, z.k1 + er > d.nz ? d.nz : z.k1 + er};
            const ZRange bw{z.k0 - eb < 0 ? 0 : z.k0 - eb, z.k1 + eb > d.nz ? d.nz : z.k1 + eb};
            JC_TRY(launch_sweep<P>(ctx, d, rw, bw, chunk, A, b, sor, *x, launches, halo ? s2 > 0 : (it > 0 || s2 > 0)));
        }
        it += (unsigned)k;
    }
    return JETCU_OK;
}

int fdm_relax_rb(jetcu_ctx* ctx, Dim3i d, ZRange z, bool halo, Mat A, const double* b, double sor, unsigned iterations,
                 double** x, double** xalt, int level) {
    if (iterations == 0 || z.k1 <= z.k0) return JETCU_OK;
    int slot = jc_prof_begin(ctx, level == 0 ? JETCU_PROF_RELAX_L0 : JETCU_PROF_RELAX_COARSE);
    uint64_t launches = 0;
    int rc = A.packed ? launch_relax<true>(ctx, d, z, halo, A, b, sor, iterations, x, xalt, &launches)
                      : launch_relax<false>(ctx, d, z, halo, A, b, sor, iterations, x, xalt, &launches);
    jc_prof_end(ctx, slot, launches, (uint64_t)d.nx * d.ny * (z.k1 - z.k0) * iterations);
    return rc;
}
int fdm_relax_rb(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double sor, unsigned iterations, double** x,
                 double** xalt, int level) {
    return fdm_relax_rb(ctx, d, zfull(d), false, A, b, sor, iterations, x, xalt, level);
}

// ---- FdmGaussSeidelSolver3::relax, src/jet/fdm_gauss_seidel_solver3.cpp:103-120: the natural (lexicographic) ordering ----------
// The smoother FdmMgSolver3 / FdmMgpcgSolver3 install unless useRedBlackOrdering is set (src/jet/fdm_mg_solver3.cpp:26-49;
// `false` is the constructors' default).  As written it is one serial loop, i fastest.  A cell reads the NEW values of its three
// lower neighbours and the OLD values of its three upper ones, so the update order only has to respect the partial order
// "lower neighbours first": every cell of a hyperplane i + j + k = h is independent of the others.  Two schedules, both with the
// per-cell body of the red-black kernels (relax_cell / relax_cell_code: the reference's expression, fdm_ops.cuh):
//  * k_relax_lex_tiles — GS_T^3 tiles, one CTA each; launch t runs, for every sweep s in flight, the tiles of tile hyperplane
//    S = t - 2 s (inside a tile the 3 GS_T - 2 cell hyperplanes run one after the other between __syncthreads).  Sweep s + 1
//    trails sweep s by two tile hyperplanes: the tiles it reads above itself (S + 1) were finished by sweep s in the previous
//    launch and are not touched again before the next one, so no tile is written and read by different CTAs of one launch.
//    n sweeps of a level with H tile hyperplanes take H + 2 (n - 1) launches instead of n H.
//  * k_relax_lex_small — levels of at most SMALL_MAX_CELLS cells: one CTA, the same pipelining at cell granularity
//    (cell hyperplane h of sweep s runs in step h + 2 s).
// Bit-identical to the serial loop (tests/test_hostcheck_linear.py replays both schedules on the CPU against the oracle,
// tests/test_gpu_linear.py compares the device with the oracle and the unmodified reference).
#define GS_T 8
#define GS_H (GS_T + 2)
// The tile's x values and a one-cell halo around them live in shared memory for the tile's 3 GS_T - 2 steps, each thread's
// coefficients, diagonal and right-hand side in registers (all global loads of a tile are issued at once, up front): a step is
// a shared-memory round trip instead of a global one.  Halo cells outside the level hold 0.0 (see relax_update).  The halo is
// read once at the start: the lower neighbour tiles carry this sweep's values from earlier launches, the upper ones the
// previous sweep's, and neither is written during this launch.
template <bool P, bool CODE>
__global__ void __launch_bounds__(GS_T* GS_T* GS_T)
    k_relax_lex_tiles(Mat A, const double* __restrict__ b, double* x, double w, int nx, int ny, int nz, int t, int s_first) {
    jc_pdl_prologue();
    __shared__ double xs[GS_H * GS_H * GS_H];
    const int ti = blockIdx.x, tj = blockIdx.y, tk = t - 2 * (s_first + (int)blockIdx.z) - ti - tj;
    const int ntz = (nz + GS_T - 1) / GS_T;
    if (tk < 0 || tk >= ntz) return;
    const int i0 = ti * GS_T, j0 = tj * GS_T, k0 = tk * GS_T;
    for (int e = threadIdx.x; e < GS_H * GS_H * GS_H; e += GS_T * GS_T * GS_T) {
        const int gi = i0 + e % GS_H - 1, gj = j0 + (e / GS_H) % GS_H - 1, gk = k0 + e / (GS_H * GS_H) - 1;
        const bool in = gi >= 0 && gi < nx && gj >= 0 && gj < ny && gk >= 0 && gk < nz;
        xs[e] = in ? x[((size_t)gk * ny + gj) * nx + gi] : 0.0;
    }
    const int cx = threadIdx.x % GS_T, cy = (threadIdx.x / GS_T) % GS_T, cz = threadIdx.x / (GS_T * GS_T);
    const int i = i0 + cx, j = j0 + cy, k = k0 + cz;
    const bool live = i < nx && j < ny && k < nz;
    const int mine = cx + cy + cz;
    const int o = ((cz + 1) * GS_H + (cy + 1)) * GS_H + (cx + 1);
    RelaxCoefs q;
    double bb = 0.0;
    if (live) {
        q = CODE ? relax_coefs_code(A, i, j, k, nx, ny, nz) : relax_coefs<P>(A, i, j, k, nx, ny, nz);
        bb = b[((size_t)k * ny + j) * nx + i];
    }
    __syncthreads();
    for (int step = 0; step <= 3 * GS_T - 3; ++step) {
        if (live && mine == step)
            xs[o] = relax_update(q, bb, w, xs[o], xs[o - 1], xs[o + 1], xs[o - GS_H], xs[o + GS_H], xs[o - GS_H * GS_H], xs[o + GS_H * GS_H]);
        __syncthreads();
    }
    if (live) x[((size_t)k * ny + j) * nx + i] = xs[o];
}

// The face-code form of the tile kernel with ONE THREAD PER x-ROW of the tile (64 threads): at step s the thread of row (cy, cz)
// updates the cell cx = s - cy - cz of its row, if there is one.  A tile hyperplane has at most 48 cells, so 512 threads per
// tile left 90 % of the resident threads idle at every step and held the SM to 4 tiles; with 64 threads, the tile's x (+ halo),
// right-hand side and code bytes in 13 KB of shared memory and the coefficients decoded from the code bytes at the step
// (relax_coefs_of_six: the bulk cell needs no load at all), 16 tiles share an SM.  The same functions on the same values as
// k_relax_lex_tiles<true, true>: bit-identical.
#define GS_C (GS_T + 1)
__global__ void __launch_bounds__(GS_T* GS_T, 16)
    k_relax_lex_rows(Mat A, const double* __restrict__ b, double* x, double w, int nx, int ny, int nz, int t, int s_first) {
    jc_pdl_prologue();
    __shared__ double xs[GS_H * GS_H * GS_H];
    __shared__ double bs[GS_T * GS_T * GS_T];
    __shared__ unsigned char cs[GS_C * GS_C * GS_C];   // the tile's codes and the layer below it on every axis
    const int ti = blockIdx.x, tj = blockIdx.y, tk = t - 2 * (s_first + (int)blockIdx.z) - ti - tj;
    const int ntz = (nz + GS_T - 1) / GS_T;
    if (tk < 0 || tk >= ntz) return;
    const int i0 = ti * GS_T, j0 = tj * GS_T, k0 = tk * GS_T;
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    for (int e = threadIdx.x; e < GS_H * GS_H * GS_H; e += GS_T * GS_T) {
        const int gi = i0 + e % GS_H - 1, gj = j0 + (e / GS_H) % GS_H - 1, gk = k0 + e / (GS_H * GS_H) - 1;
        const bool in = gi >= 0 && gi < nx && gj >= 0 && gj < ny && gk >= 0 && gk < nz;
        xs[e] = in ? x[(size_t)gk * sz + (size_t)gj * sy + gi] : 0.0;
    }
    for (int e = threadIdx.x; e < GS_T * GS_T * GS_T; e += GS_T * GS_T) {
        const int gi = i0 + e % GS_T, gj = j0 + (e / GS_T) % GS_T, gk = k0 + e / (GS_T * GS_T);
        bs[e] = (gi < nx && gj < ny && gk < nz) ? b[(size_t)gk * sz + (size_t)gj * sy + gi] : 0.0;
    }
    for (int e = threadIdx.x; e < GS_C * GS_C * GS_C; e += GS_T * GS_T) {
        const int gi = i0 + e % GS_C - 1, gj = j0 + (e / GS_C) % GS_C - 1, gk = k0 + e / (GS_C * GS_C) - 1;
        const bool in = gi >= 0 && gi < nx && gj >= 0 && gj < ny && gk >= 0 && gk < nz;
        cs[e] = in ? jc_ldg(A.code + (size_t)gk * sz + (size_t)gj * sy + gi) : (unsigned char)0;
    }
    const int cy = threadIdx.x % GS_T, cz = threadIdx.x / GS_T;
    const int j = j0 + cy, k = k0 + cz;
    const bool row_live = j < ny && k < nz;
    __syncthreads();
    for (int step = 0; step <= 3 * GS_T - 3; ++step) {
        const int cx = step - cy - cz;
        if (row_live && cx >= 0 && cx < GS_T && i0 + cx < nx) {
            const int o = ((cz + 1) * GS_H + (cy + 1)) * GS_H + (cx + 1);
            const int oc = ((cz + 1) * GS_C + (cy + 1)) * GS_C + (cx + 1);
            const unsigned six = relax_six(cs[oc], cs[oc - 1], cs[oc - GS_C], cs[oc - GS_C * GS_C]);
            const RelaxCoefs q = relax_coefs_of_six(A, six, (size_t)k * sz + (size_t)j * sy + (size_t)(i0 + cx), sy, sz);
            xs[o] = relax_update(q, bs[(cz * GS_T + cy) * GS_T + cx], w, xs[o], xs[o - 1], xs[o + 1], xs[o - GS_H], xs[o + GS_H],
                                 xs[o - GS_H * GS_H], xs[o + GS_H * GS_H]);
        }
        __syncthreads();
    }
    if (row_live) {
        double* dst = x + (size_t)k * sz + (size_t)j * sy + i0;
        const int o = ((cz + 1) * GS_H + (cy + 1)) * GS_H + 1;
#pragma unroll
        for (int cx = 0; cx < GS_T; ++cx)
            if (i0 + cx < nx) dst[cx] = xs[o + cx];
    }
}

// one CTA, whole level; SMEM: x in shared memory for the duration (as k_relax_small)
template <bool P, bool SMEM>
__global__ void __launch_bounds__(SMALL_THREADS) k_relax_lex_small(Mat A, const double* __restrict__ b, double* x, double w,
                                                                   int nx, int ny, int nz, unsigned iterations) {
    __shared__ double xs[SMEM ? SMALL_SMEM_CELLS : 1];
    const int n = nx * ny * nz;
    const int steps = (nx + ny + nz - 2) + 2 * ((int)iterations - 1);
    double* xw = x;
    if (SMEM) {
        for (int c = threadIdx.x; c < n; c += SMALL_THREADS) xs[c] = x[c];
        __syncthreads();
        xw = xs;
    }
    for (int step = 0; step < steps; ++step) {
        for (int c = threadIdx.x; c < n; c += SMALL_THREADS) {
            const int i = c % nx, j = (c / nx) % ny, k = c / (nx * ny);
            const int lag = step - (i + j + k);   // = 2 x (the sweep this cell is in at this step)
            if (lag < 0 || (lag & 1) || lag / 2 >= (int)iterations) continue;
            relax_cell<P>(A, b, xw, w, i, j, k, nx, ny, nz);
        }
        __syncthreads();
    }
    if (SMEM)
        for (int c = threadIdx.x; c < n; c += SMALL_THREADS) x[c] = xs[c];
}

// sweeps in flight per group of launches (gridDim.z); a group of n sweeps costs H + 2 (n - 1) launches
#define GS_MAX_INFLIGHT 32
template <bool P>
static int launch_relax_lex(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double sor, unsigned iterations, double* x,
                            uint64_t* launches) {
    *launches = 0;
    if (d.n() <= SMALL_MAX_CELLS) {
        if (d.n() <= SMALL_SMEM_CELLS) JC_LAUNCH(ctx, (k_relax_lex_small<P, true>), 1, SMALL_THREADS, 0, A, b, x, sor, d.nx, d.ny, d.nz, iterations);
        else JC_LAUNCH(ctx, (k_relax_lex_small<P, false>), 1, SMALL_THREADS, 0, A, b, x, sor, d.nx, d.ny, d.nz, iterations);
        *launches = 1;
        return JETCU_OK;
    }
    const int ntx = (d.nx + GS_T - 1) / GS_T, nty = (d.ny + GS_T - 1) / GS_T, ntz = (d.nz + GS_T - 1) / GS_T;
    const int last = ntx + nty + ntz - 3;   // highest tile hyperplane
    const bool code = P && A.code != nullptr;
    for (unsigned done = 0; done < iterations;) {
        const int n = (int)((iterations - done < GS_MAX_INFLIGHT) ? iterations - done : GS_MAX_INFLIGHT);
        for (int t = 0; t <= last + 2 * (n - 1); ++t) {
            // sweeps s with 0 <= t - 2 s <= last
            int s_lo = t - last > 0 ? (t - last + 1) / 2 : 0, s_hi = t / 2 < n - 1 ? t / 2 : n - 1;
            if (s_lo > s_hi) continue;
            dim3 grid(ntx, nty, s_hi - s_lo + 1);
            const bool after = t > 0;   // the previous operation of the stream is the previous launch of this group
            // one thread per tile row pays where a launch has enough tiles to fill 16 per SM (GPU call 22: 256^3 80.7 -> 68.6 ms
            // per solve, but 128^3 18.8 -> 27.7: few tiles per launch, and the 64-thread prologue is the longer one); 1 / 0 force it
            static const int rows_env = getenv("JETCU_LEX_ROWS") ? atoi(getenv("JETCU_LEX_ROWS")) : -1;
            const bool rows = rows_env >= 0 ? rows_env != 0 : (size_t)ntx * nty * ntz >= 16384;
            if (code && rows) JC_LAUNCH_PDL(ctx, after, k_relax_lex_rows, grid, GS_T * GS_T, 0, A, b, x, sor, d.nx, d.ny, d.nz, t, s_lo);
            else if (code) JC_LAUNCH_PDL(ctx, after, (k_relax_lex_tiles<P, true>), grid, GS_T * GS_T * GS_T, 0, A, b, x, sor, d.nx, d.ny, d.nz, t, s_lo);
            else JC_LAUNCH_PDL(ctx, after, (k_relax_lex_tiles<P, false>), grid, GS_T * GS_T * GS_T, 0, A, b, x, sor, d.nx, d.ny, d.nz, t, s_lo);
            *launches += 1;
        }
        done += (unsigned)n;
    }
    return JETCU_OK;
}

int fdm_relax_lex(jetcu_ctx* ctx, Dim3i d, Mat A, const double* b, double sor, unsigned iterations, double* x, int level) {
    if (iterations == 0 || d.n() == 0) return JETCU_OK;
    if (jc_comm_world(ctx) > 1)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "lexicographic Gauss-Seidel (useRedBlackOrdering = false) is not slab-"
                          "distributed: use useRedBlackOrdering = true on a multi-GPU communicator");
    int slot = jc_prof_begin(ctx, level == 0 ? JETCU_PROF_RELAX_L0 : JETCU_PROF_RELAX_COARSE);
    uint64_t launches = 0;
    int rc = A.packed ? launch_relax_lex<true>(ctx, d, A, b, sor, iterations, x, &launches)
                      : launch_relax_lex<false>(ctx, d, A, b, sor, iterations, x, &launches);
    jc_prof_end(ctx, slot, launches, (uint64_t)d.n() * iterations);
    return rc;
}

__global__ void __launch_bounds__(BX* BY) k_restrict(const double* __restrict__ finer, double* __restrict__ coarser,
                                                     int cnx, int cny, int cnz, int k0) {
    int i = blockIdx.x * BX + threadIdx.x, j = blockIdx.y * BY + threadIdx.y, k = k0 + blockIdx.z;
    if (i >= cnx || j >= cny) return;
    restrict_cell(finer, coarser, i, j, k, cnx, cny, cnz);
}

__global__ void __launch_bounds__(BX* BY) k_correct(const double* __restrict__ coarser, double* __restrict__ finer,
                                                    int nx, int ny, int nz, int k0) {
    int i = blockIdx.x * BX + threadIdx.x, j = blockIdx.y * BY + threadIdx.y, k = k0 + blockIdx.z;
    if (i >= nx || j >= ny) return;
    correct_cell(coarser, finer, i, j, k, nx, ny, nz);
}

// `cz` = the COARSE planes to produce; reads fine planes 2*k0-1 .. 2*k1 (clamped to the level).
int fdm_restrict(jetcu_ctx* ctx, Dim3i fine, ZRange cz, const double* finer, double* coarser) {
    Dim3i c{fine.nx / 2, fine.ny / 2, fine.nz / 2};
    if (cz.k1 <= cz.k0) return JETCU_OK;
    JC_LAUNCH(ctx, k_restrict, grid3z(c, cz), dim3(BX, BY), 0, finer, coarser, c.nx, c.ny, c.nz, cz.k0);
    return JETCU_OK;
}
int fdm_restrict(jetcu_ctx* ctx, Dim3i fine, const double* finer, double* coarser) {
    return fdm_restrict(ctx, fine, ZRange{0, fine.nz / 2}, finer, coarser);
}

// `z` = the FINE planes to correct; reads coarse planes k0/2-1 .. (k1-1)/2+1 (clamped).
int fdm_correct(jetcu_ctx* ctx, Dim3i fine, ZRange z, const double* coarser, double* finer) {
    if (z.k1 <= z.k0) return JETCU_OK;
    JC_LAUNCH(ctx, k_correct, grid3z(fine, z), dim3(BX, BY), 0, coarser, finer, fine.nx, fine.ny, fine.nz, z.k0);
    return JETCU_OK;
}
int fdm_correct(jetcu_ctx* ctx, Dim3i fine, const double* coarser, double* finer) {
    return fdm_correct(ctx, fine, zfull(fine), coarser, finer);
}

// ------------------------------------------------------------------------------------------------ BLAS-1
#define DOT_BLOCK 256
#define DOT_GRID 592  // 4 CTAs per SM x 148 SMs; fixed so the reduction tree (and the result bits) never change

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double s[DOT_BLOCK / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < DOT_BLOCK / 32) ? s[threadIdx.x] : 0.0;
        r = warp_sum(r);
    }
    return r;  // valid in thread 0
}

__global__ void __launch_bounds__(DOT_BLOCK) k_dot_partial(const double* __restrict__ a, const double* __restrict__ b,
                                                           double* __restrict__ partials, size_t n) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * DOT_BLOCK + threadIdx.x; i < n; i += (size_t)DOT_GRID * DOT_BLOCK)
        acc += a[i] * b[i];
    double r = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

__global__ void __launch_bounds__(DOT_BLOCK) k_dot_final(const double* __restrict__ partials, int count,
                                                         double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += DOT_BLOCK) acc += partials[i];
    double r = block_sum(acc);
    if (threadIdx.x == 0) out[0] = r;
}

int fdm_dot(jetcu_ctx* ctx, size_t n, const double* a, const double* b, double* d_result) {
    JC_LAUNCH(ctx, k_dot_partial, DOT_GRID, DOT_BLOCK, 0, a, b, ctx->d_partials, n);
    JC_LAUNCH(ctx, k_dot_final, 1, DOT_BLOCK, 0, ctx->d_partials, DOT_GRID, d_result);
    return JETCU_OK;
}

// Dot product whose bits do not depend on how the grid is split into z-slabs: one partial per (z-plane, segment),
// each summed by one CTA in a fixed order; the partials of all planes (all ranks) are then reduced by one CTA in a
// fixed tree.  Used by PCG, so that an N-GPU solve takes exactly the iterations of the 1-GPU solve.
__global__ void __launch_bounds__(DOT_BLOCK) k_dot_plane_partial(const double* __restrict__ a, const double* __restrict__ b,
                                                                 double* __restrict__ partials, size_t plane, int seg,
                                                                 int k0) {
    const int k = k0 + blockIdx.y, s = blockIdx.x;
    const size_t chunk = (plane + seg - 1) / seg;
    const size_t lo = (size_t)s * chunk, hi = (lo + chunk < plane) ? lo + chunk : plane;
    const size_t base = (size_t)k * plane;
    double acc = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += DOT_BLOCK) acc += a[base + i] * b[base + i];
    double r = block_sum(acc);
    if (threadIdx.x == 0) partials[(size_t)k * seg + s] = r;
}

static inline int dot_segments(Dim3i d) {
    int seg = JC_PARTIALS_CAP / d.nz;
    return seg > 4 ? 4 : seg;
}

int fdm_dot_planes(jetcu_ctx* ctx, Dim3i d, ZRange z, bool dist, const double* a, const double* b, double* d_result) {
    const int seg = dot_segments(d);
    if (seg < 1) return jetcu_fail(ctx, JETCU_ERR_INVALID_ARG, "nz = %d exceeds the dot-product partial buffer", d.nz);
    const size_t plane = (size_t)d.nx * d.ny;
    if (z.k1 > z.k0)
        JC_LAUNCH(ctx, k_dot_plane_partial, dim3(seg, z.k1 - z.k0), DOT_BLOCK, 0, a, b, ctx->d_partials, plane, seg, z.k0);
    if (dist) JC_TRY(comm_allgather_inplace(ctx, ctx->d_partials, (size_t)(z.k1 - z.k0) * seg * sizeof(double)));
    JC_LAUNCH(ctx, k_dot_final, 1, DOT_BLOCK, 0, ctx->d_partials, d.nz * seg, d_result);
    return JETCU_OK;
}

// FdmBlas3::axpy, src/jet/fdm_linear_system3.cpp:67-77.
__global__ void k_axpy(double a, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out,
                       size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + y[i];
}

__global__ void k_pcg_update(double alpha, const double* __restrict__ d, const double* __restrict__ q,
                             double* __restrict__ x, double* __restrict__ r, int update_r, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = alpha * d[i] + x[i];
    if (update_r) r[i] = -alpha * q[i] + r[i];
}

// the same update with alpha = sigma / (d . q) formed on the device from the two dot-product results (the IEEE division the host
// would do: same bits), so that the solver does not have to wait for d . q on the host
__global__ void k_pcg_update_dev(const double* __restrict__ sigma, const double* __restrict__ dq, const double* __restrict__ d,
                                 const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, int update_r, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double alpha = *sigma / *dq;
    x[i] = alpha * d[i] + x[i];
    if (update_r) r[i] = -alpha * q[i] + r[i];
}

int fdm_axpy(jetcu_ctx* ctx, size_t n, double a, const double* x, const double* y, double* out) {
    JC_LAUNCH(ctx, k_axpy, jc_blocks(n, 256), 256, 0, a, x, y, out, n);
    return JETCU_OK;
}

int fdm_pcg_update(jetcu_ctx* ctx, size_t n, double alpha, const double* d, const double* q, double* x, double* r,
                   bool update_r) {
    JC_LAUNCH(ctx, k_pcg_update, jc_blocks(n, 256), 256, 0, alpha, d, q, x, r, update_r ? 1 : 0, n);
    return JETCU_OK;
}

// ------------------------------------------------------------------------------------------------ MG + PCG
int mg_alloc_vectors(jetcu_ctx* ctx, MgLevels* mg) {
    for (int l = 0; l < mg->levels; ++l) {
        size_t n = mg->dims[l].n();
        for (int q = 0; q < 4; ++q) JC_TRY(jc_alloc(ctx, &mg->own_vec[l][q], n));
        mg->x[l] = mg->own_vec[l][0]; mg->xalt[l] = mg->own_vec[l][1];
        mg->b[l] = mg->own_vec[l][2]; mg->buf[l] = mg->own_vec[l][3];
    }
    return mg_make_plan(ctx, mg);
}

void mg_free(jetcu_ctx* ctx, MgLevels* mg) {
    for (auto& g : mg->vgraph) {
        if (g.exec) { cudaStreamSynchronize(ctx->stream); cudaGraphExecDestroy(g.exec); ctx->live_graphs--; }
        g = MgLevels::VGraph();
    }
    jc_free(ctx, mg->pcg_d); jc_free(ctx, mg->pcg_s);
    for (int l = 0; l < JC_MAX_LEVELS; ++l) {
        for (int q = 0; q < 4; ++q) jc_free(ctx, mg->own_vec[l][q]);
        mg->x[l] = mg->xalt[l] = mg->b[l] = mg->buf[l] = nullptr;
    }
    mg->levels = 0;
}

// z-slab plan of an MG hierarchy over the ctx's communicator.  Level l is slab-distributed while every
// rank keeps an even number (>= 4) of planes and the level is too big for the single-CTA smoother; all coarser levels
// are REPLICATED: their right-hand side is all-gathered once on the way down and every rank relaxes the whole (tiny)
// level redundantly, which costs no further messages and keeps the arithmetic identical to one GPU.
// Coarser levels than this many cells are replicated rather than slab-distributed (every result is bit-identical either
// way).  Default 300 000 = "64^3 and below": at those sizes a colour pass is launch-latency-bound whether it covers the
// whole level or an eighth of it (a few microseconds), while a distributed level visit costs 10-14 dependent neighbour
// exchanges of ~27 us each (measured) — at 8 slabs of 256^3 that removes 28 of the ~54 exchanges of a
// V-cycle.  128^3 (2.1 M cells) stays distributed: its sweeps are bandwidth-bound, so splitting them still pays.
// JETCU_SLAB_MIN_CELLS overrides the threshold for A/B runs (8192 = the round-1 plan: distribute down to 4 planes per rank).
static size_t slab_min_cells() {
    static long long v = -1;
    if (v < 0) {
        const char* e = getenv("JETCU_SLAB_MIN_CELLS");
        v = e ? atoll(e) : 300000ll;
        if (v < (long long)SMALL_MAX_CELLS) v = (long long)SMALL_MAX_CELLS;
    }
    return (size_t)v;
}

int slab_plan_compute(const Dim3i* dims, int levels, int world, int rank, SlabPlan* out) {
    SlabPlan& pl = *out;
    pl = SlabPlan();
    pl.world = world; pl.rank = rank;
    for (int l = 0; l < JC_MAX_LEVELS; ++l) { pl.z0[l] = 0; pl.z1[l] = l < levels ? dims[l].nz : 0; }
    if (world <= 1) return 0;
    int l = 0;
    for (; l < levels; ++l) {
        const int nz = dims[l].nz;
        if (nz % (2 * world) != 0 || nz / world < 4 || dims[l].n() <= (l == 0 ? (size_t)SMALL_MAX_CELLS : slab_min_cells())) break;
        pl.z0[l] = rank * (nz / world); pl.z1[l] = (rank + 1) * (nz / world);
    }
    pl.dist_levels = l;
    return l == 0 ? -1 : 0;
}

int mg_make_plan(jetcu_ctx* ctx, MgLevels* mg) {
    if (slab_plan_compute(mg->dims, mg->levels, jc_comm_world(ctx), jc_comm_rank(ctx), &mg->plan) != 0)
        return jetcu_fail(ctx, JETCU_ERR_INVALID_ARG, "z-slab decomposition needs nz (%d) to be a multiple of 2 x %d ranks "
                          "with at least 4 planes per rank", mg->dims[0].nz, jc_comm_world(ctx));
    return JETCU_OK;
}

extern "C" int jetcu_slab_plan(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, int world, int rank,
                               uint32_t* dist_levels, uint64_t* z0, uint64_t* z1) {
    if (levels < 1 || levels > JC_MAX_LEVELS || world < 1 || rank < 0 || rank >= world) return JETCU_ERR_INVALID_ARG;
    Dim3i dims[JC_MAX_LEVELS];
    for (uint32_t l = 0; l < levels; ++l) dims[l] = Dim3i{(int)(nx >> l), (int)(ny >> l), (int)(nz >> l)};
    SlabPlan pl;
    if (slab_plan_compute(dims, (int)levels, world, rank, &pl) != 0) return JETCU_ERR_INVALID_ARG;
    if (dist_levels) *dist_levels = (uint32_t)pl.dist_levels;
    for (uint32_t l = 0; l < levels; ++l) {
        if (z0) z0[l] = (uint64_t)pl.z0[l];
        if (z1) z1[l] = (uint64_t)pl.z1[l];
    }
    return JETCU_OK;
}

// internal::mgVCycle, include/jet/detail/mg-inl.h:16-68, with FdmMgSolver3's relax lambdas
// (src/jet/fdm_mg_solver3.cpp:26-49).  The residual + l2Norm the reference evaluates on the way up only feed
// MgResult.lastResidualNorm, which FdmMgpcgSolver3::Preconditioner discards; they never touch x and are skipped.
// mg->x[l] / mg->xalt[l] may be exchanged by the smoother; mg->x[l] always names the current iterate.
// On a slab-distributed level every rank works on its planes [z0, z1) of the full-size (globally indexed) arrays and
// refreshes the two planes around them from its neighbours before each stencil application.
// the smoother FdmMgSolver3's constructor installs (src/jet/fdm_mg_solver3.cpp:26-49): relaxRedBlack or the natural relax
static int vc_relax(jetcu_ctx* ctx, MgLevels* mg, const jetcu_mgpcg_params* p, int l, bool dist, ZRange z, unsigned iterations) {
    if (p->use_red_black)
        return fdm_relax_rb(ctx, mg->dims[l], z, dist, mg->A[l], mg->b[l], p->sor_factor, iterations, &mg->x[l], &mg->xalt[l], l);
    return fdm_relax_lex(ctx, mg->dims[l], mg->A[l], mg->b[l], p->sor_factor, iterations, mg->x[l], l);
}

int mg_vcycle(jetcu_ctx* ctx, MgLevels* mg, const jetcu_mgpcg_params* p, int l) {
    const Dim3i d = mg->dims[l];
    const SlabPlan& pl = mg->plan;
    const bool dist = l < pl.dist_levels;
    const ZRange z{pl.z0[l], pl.z1[l]};
    const size_t plane = (size_t)d.nx * d.ny * sizeof(double);
    // the smoother's redundant halo cells read b next to the slab; b[l] does not change during this level's visit
    if (dist && fdm_relax_rhs_halo(z.k1 - z.k0, (size_t)d.nx * d.ny) > 0)
        JC_TRY(comm_halo_planes(ctx, mg->b[l], plane, z.k0, z.k1, d.nz, fdm_relax_rhs_halo(z.k1 - z.k0, (size_t)d.nx * d.ny)));
    JC_TRY(vc_relax(ctx, mg, p, l, dist, z, p->n_restriction));
    if (l + 1 < mg->levels) {
        const Dim3i dc = mg->dims[l + 1];
        const ZRange cz{z.k0 / 2, z.k1 / 2};
        int slot = jc_prof_begin(ctx, JETCU_PROF_TRANSFER);
        if (dist) JC_TRY(comm_halo_planes(ctx, mg->x[l], plane, z.k0, z.k1, d.nz));
        JC_TRY(fdm_residual(ctx, d, z, mg->A[l], mg->x[l], mg->b[l], mg->buf[l]));
        if (dist) JC_TRY(comm_halo_planes(ctx, mg->buf[l], plane, z.k0, z.k1, d.nz));
        JC_TRY(fdm_restrict(ctx, d, cz, mg->buf[l], mg->b[l + 1]));
        if (dist && l + 1 >= pl.dist_levels)  // first replicated level: every rank needs the whole right-hand side
            JC_TRY(comm_allgather_inplace(ctx, mg->b[l + 1], (size_t)(cz.k1 - cz.k0) * dc.nx * dc.ny * sizeof(double)));
        JC_TRY(jc_zero(ctx, mg->x[l + 1], dc.n()));
        jc_prof_end(ctx, slot, 2, (uint64_t)d.nx * d.ny * (z.k1 - z.k0));
        JC_TRY(mg_vcycle(ctx, mg, p, l + 1));
        slot = jc_prof_begin(ctx, JETCU_PROF_TRANSFER);
        if (l + 1 < pl.dist_levels)
            JC_TRY(comm_halo_planes(ctx, mg->x[l + 1], (size_t)dc.nx * dc.ny * sizeof(double), pl.z0[l + 1], pl.z1[l + 1], dc.nz));
        JC_TRY(fdm_correct(ctx, d, z, mg->x[l + 1], mg->x[l]));
        jc_prof_end(ctx, slot, 1, (uint64_t)d.nx * d.ny * (z.k1 - z.k0));
        JC_TRY(vc_relax(ctx, mg, p, l, dist, z, l > 0 ? p->n_correction : p->n_final));
    } else {
        JC_TRY(vc_relax(ctx, mg, p, l, dist, z, p->n_coarsest));
    }
    return JETCU_OK;
}

// FdmMgpcgSolver3::Preconditioner::solve, src/jet/fdm_mgpcg_solver3.cpp:21-36: the V-cycle's initial guess is the
// CURRENT content of the output vector.  We run the cycle directly on (b_in, *x_inout) as level 0's b and x, with
// *scratch as the ping-pong partner; on return *x_inout names the buffer holding the result (the two pointers may
// have been exchanged).
static int precondition_eager(jetcu_ctx* ctx, MgLevels* mg, const jetcu_mgpcg_params* p, double* b_in, double** x_inout,
                              double** scratch) {
    double *saved_x = mg->x[0], *saved_alt = mg->xalt[0], *saved_b = mg->b[0];
    mg->x[0] = *x_inout; mg->xalt[0] = *scratch; mg->b[0] = b_in;
    int rc = mg_vcycle(ctx, mg, p, 0);
    *x_inout = mg->x[0]; *scratch = mg->xalt[0];
    mg->x[0] = saved_x; mg->xalt[0] = saved_alt; mg->b[0] = saved_b;
    return rc;
}

// FNV-1a over everything a captured V-cycle bakes in besides (b, x): matrices (passed by value), vectors, sweeps, plan.
static uint64_t vcycle_key(const MgLevels* mg, const jetcu_mgpcg_params* p) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void* d, size_t n) {
        const unsigned char* c = static_cast<const unsigned char*>(d);
        for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
    };
    mix(&mg->levels, sizeof(int));
    for (int l = 0; l < mg->levels; ++l) {
        const Mat& A = mg->A[l];
        const void* ptrs[] = {A.center, A.right, A.up, A.front, A.cr, A.cu, A.cf, A.code, mg->x[l], mg->b[l], mg->buf[l]};
        mix(ptrs, sizeof(ptrs));
        const double sc[] = {A.sx, A.sy, A.sz, A.c6};
        mix(sc, sizeof(sc));
        mix(&mg->dims[l], sizeof(Dim3i));
        mix(&mg->plan.z0[l], sizeof(int)); mix(&mg->plan.z1[l], sizeof(int));
    }
    mix(&mg->plan.world, sizeof(int)); mix(&mg->plan.rank, sizeof(int)); mix(&mg->plan.dist_levels, sizeof(int));
    const uint32_t it[] = {p->n_restriction, p->n_correction, p->n_coarsest, p->n_final, (uint32_t)(p->use_red_black != 0)};
    mix(it, sizeof(it));
    mix(&p->sor_factor, sizeof(double));
    return h;
}

// Replays (or first captures) the V-cycle on (b_in, *x_inout) as a CUDA graph.  The smoother works in place, so the
// pointers never change hands; NCCL halo exchanges are captured along with the kernels.  Falls back to eager launches
// while kernel-family profiling is on (its event pairs cannot live inside a graph) or once a capture has failed.
static int precondition(jetcu_ctx* ctx, MgLevels* mg, const jetcu_mgpcg_params* p, double* b_in, double** x_inout,
                        double** scratch) {
    if (!ctx->graphs_ok || ctx->profile) return precondition_eager(ctx, mg, p, b_in, x_inout, scratch);
    // level 0's x / b slots are overridden per call, keep them out of the key
    double *saved_x = mg->x[0], *saved_b = mg->b[0];
    mg->x[0] = nullptr; mg->b[0] = nullptr;
    const uint64_t key = vcycle_key(mg, p);
    mg->x[0] = saved_x; mg->b[0] = saved_b;
    MgLevels::VGraph* slot = nullptr;
    for (auto& g : mg->vgraph)
        if (g.exec && g.b == b_in && g.x == *x_inout && g.key == key) slot = &g;
    if (!slot) {
        slot = (mg->vgraph[0].stamp <= mg->vgraph[1].stamp) ? &mg->vgraph[0] : &mg->vgraph[1];
        if (slot->exec) { cudaStreamSynchronize(ctx->stream); cudaGraphExecDestroy(slot->exec); ctx->live_graphs--; }
        *slot = MgLevels::VGraph();
        uint64_t m0, b0, m1, b1;
        comm_counters(ctx, &m0, &b0);
        const uint64_t l0 = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            ctx->graphs_ok = false;
            return precondition_eager(ctx, mg, p, b_in, x_inout, scratch);
        }
        double *x = *x_inout, *sc = *scratch;
        int rc = precondition_eager(ctx, mg, p, b_in, &x, &sc);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        cudaGraphExec_t exec = nullptr;
        if (rc == JETCU_OK && e == cudaSuccess && graph && x == *x_inout) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        comm_counters(ctx, &m1, &b1);
        if (rc != JETCU_OK || e != cudaSuccess || !exec) {  // nothing ran: counters back, eager from now on
            cudaGetLastError();
            ctx->launches = l0;
            comm_counters_add(ctx, m0 - m1, b0 - b1);
            ctx->graphs_ok = false;
            return precondition_eager(ctx, mg, p, b_in, x_inout, scratch);
        }
        slot->exec = exec; slot->b = b_in; slot->x = *x_inout; slot->key = key;
        ctx->live_graphs++;
        slot->launches = ctx->launches - l0; slot->messages = m1 - m0; slot->bytes = b1 - b0;
        ctx->launches = l0;
        comm_counters_add(ctx, m0 - m1, b0 - b1);  // the capture itself sent nothing
    }
    slot->stamp = ++mg->vgraph_clock;
    JC_CUDA(ctx, cudaGraphLaunch(slot->exec, ctx->stream));
    ctx->launches += slot->launches;
    comm_counters_add(ctx, slot->messages, slot->bytes);
    return JETCU_OK;
}

static int read_scalar(jetcu_ctx* ctx, const double* d, double* out) {
    JC_TRY(jc_d2h(ctx, ctx->h_scalars, d, 1));
    JC_TRY(jc_sync(ctx));
    *out = ctx->h_scalars[0];
    return JETCU_OK;
}

// pcg<FdmBlas3, Preconditioner>, include/jet/detail/cg-inl.h:18-94, as called by FdmMgpcgSolver3::solve
// (src/jet/fdm_mgpcg_solver3.cpp:54-80): x = 0 first; exit test sigma > tol^2 (also leaves on sigma <= 0);
// residual recomputed when iter % 50 == 0 (iter > 0) or after sigma rose.
// With a multi-rank communicator on the ctx the solve is z-slab distributed (mg->plan): b must be valid on this rank's
// planes, x is complete (all-gathered) on return, and — thanks to fdm_dot_planes — every rank sees the same scalars,
// takes the same branches and produces the bits of the single-GPU solve.
int mgpcg_solve_device(jetcu_ctx* ctx, MgLevels* mg, const double* b, double* x, const jetcu_mgpcg_params* p,
                       jetcu_mgpcg_result* result) {
    if (!p->use_red_black && jc_comm_world(ctx) > 1)
        return jetcu_fail(ctx, JETCU_ERR_UNSUPPORTED, "lexicographic Gauss-Seidel (useRedBlackOrdering = false) is not slab-"
                          "distributed: use useRedBlackOrdering = true on a multi-GPU communicator");
    JC_TRY(mg_make_plan(ctx, mg));
    const Dim3i d = mg->dims[0];
    const size_t n = d.n();
    const SlabPlan& pl = mg->plan;
    const bool dist = pl.dist_levels > 0;
    const ZRange z{pl.z0[0], pl.z1[0]};
    const size_t nxy = (size_t)d.nx * d.ny, off = (size_t)z.k0 * nxy, cnt = (size_t)(z.k1 - z.k0) * nxy;
    const size_t plane = nxy * sizeof(double);
    // level 0's own x / xalt / b are idle during PCG (the V-cycle runs on the PCG vectors): reuse them as r, q and the
    // ping-pong scratch t; buf[0] stays the residual temporary of the V-cycle.
    double *r = mg->own_vec[0][0], *q = mg->own_vec[0][1], *t = mg->own_vec[0][2];
    if (!mg->pcg_d) JC_TRY(jc_alloc(ctx, &mg->pcg_d, n));
    if (!mg->pcg_s) JC_TRY(jc_alloc(ctx, &mg->pcg_s, n));
    double *dd = mg->pcg_d, *s = mg->pcg_s;
    cudaEvent_t e0, e1;
    JC_CUDA(ctx, cudaEventCreate(&e0)); JC_CUDA(ctx, cudaEventCreate(&e1));
    JC_CUDA(ctx, cudaEventRecord(e0, ctx->stream));

    JC_TRY(jc_zero(ctx, x, n)); JC_TRY(jc_zero(ctx, dd, n)); JC_TRY(jc_zero(ctx, q, n)); JC_TRY(jc_zero(ctx, s, n));
    JC_TRY(fdm_residual(ctx, d, z, mg->A[0], x, b, r));  // x = 0 everywhere: no halo needed
    JC_TRY(precondition(ctx, mg, p, r, &dd, &t));
    double sigmaNew;
    JC_TRY(fdm_dot_planes(ctx, d, z, dist, r, dd, ctx->d_scalars));
    JC_TRY(read_scalar(ctx, ctx->d_scalars, &sigmaNew));

    unsigned iter = 0;
    bool trigger = false;
    while (sigmaNew > p->tolerance * p->tolerance && iter < p->max_iterations) {
        {
            int slot = jc_prof_begin(ctx, JETCU_PROF_MVM_L0);
            if (dist) JC_TRY(comm_halo_planes(ctx, dd, plane, z.k0, z.k1, d.nz));
            JC_TRY(fdm_mvm(ctx, d, z, mg->A[0], dd, q));
            jc_prof_end(ctx, slot, 1, cnt);
        }
        // alpha = sigmaNew / (d . q) is formed on the device (d_scalars[0] still holds sigmaNew, the last r . s): one host round
        // trip per iteration (sigma, for the loop control) instead of two
        JC_TRY(fdm_dot_planes(ctx, d, z, dist, dd, q, ctx->d_scalars + 1));
        bool recompute = trigger || (iter % 50 == 0 && iter > 0);
        JC_LAUNCH(ctx, k_pcg_update_dev, jc_blocks(cnt, 256), 256, 0, ctx->d_scalars, ctx->d_scalars + 1, dd + off, q + off, x + off,
                  r + off, recompute ? 0 : 1, cnt);
        if (recompute) {
            if (dist) JC_TRY(comm_halo_planes(ctx, x, plane, z.k0, z.k1, d.nz));
            JC_TRY(fdm_residual(ctx, d, z, mg->A[0], x, b, r));
            trigger = false;
        }
        JC_TRY(precondition(ctx, mg, p, r, &s, &t));
        double sigmaOld = sigmaNew;
        JC_TRY(fdm_dot_planes(ctx, d, z, dist, r, s, ctx->d_scalars));
        JC_TRY(read_scalar(ctx, ctx->d_scalars, &sigmaNew));
        if (sigmaNew > sigmaOld) trigger = true;
        double beta = sigmaNew / sigmaOld;
        JC_TRY(fdm_axpy(ctx, cnt, beta, dd + off, s + off, dd + off));
        ++iter;
    }
    if (dist) JC_TRY(comm_allgather_inplace(ctx, x, cnt * sizeof(double)));  // pressure complete on every rank
    if (dist) JC_TRY(comm_p2p_check(ctx));
    JC_CUDA(ctx, cudaEventRecord(e1, ctx->stream));

    if (result) {
        result->iterations = iter;
        result->last_residual = sqrt(fabs(sigmaNew));
        result->returned = (result->last_residual <= p->tolerance || iter < p->max_iterations) ? 1 : 0;
        double rr, bb;
        JC_TRY(fdm_residual(ctx, d, z, mg->A[0], x, b, r));  // x complete after the all-gather: halos valid
        JC_TRY(fdm_dot_planes(ctx, d, z, dist, r, r, ctx->d_scalars));
        JC_TRY(fdm_dot_planes(ctx, d, z, dist, b, b, ctx->d_scalars + 1));
        JC_TRY(jc_d2h(ctx, ctx->h_scalars, ctx->d_scalars, 2));
        JC_TRY(jc_sync(ctx));
        rr = ctx->h_scalars[0]; bb = ctx->h_scalars[1];
        result->true_residual_l2 = sqrt(rr);
        result->b_l2 = sqrt(bb);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        result->solve_ms = ms;
    } else {
        JC_TRY(jc_sync(ctx));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    // dd / s / t may have been exchanged by a ping-pong smoother (the current one works in place: they never are)
    mg->own_vec[0][2] = t; mg->pcg_d = dd; mg->pcg_s = s;
    mg->x[0] = mg->own_vec[0][0]; mg->xalt[0] = mg->own_vec[0][1]; mg->b[0] = mg->own_vec[0][2];
    return JETCU_OK;
}
