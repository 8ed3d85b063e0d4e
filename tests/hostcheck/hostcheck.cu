// hostcheck.cu — TEST INFRASTRUCTURE.  Runs the product's own __host__ __device__ kernel bodies
// (fluid-engine-dev_b200/csrc/collider.cuh, collider_ops.cuh, sampling.cuh) on the CPU, with plain loops in place of
// the kernel launches, so that `-m "not gpu"` tests can compare the CUDA source bit-for-bit with the oracle where no
// GPU exists.  Built by tests/hostcheck/build.py with nvcc (host code only; no device is touched).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <algorithm>
#include <numeric>

#include "../../fluid-engine-dev_b200/csrc/fdm_ops.cuh"
#include "../../fluid-engine-dev_b200/csrc/pressure_ops.cuh"
#include "../../fluid-engine-dev_b200/csrc/p2g_ops.cuh"
#include "../../fluid-engine-dev_b200/csrc/iccg_ops.cuh"
#include "../../fluid-engine-dev_b200/csrc/smoke_ops.cuh"
#include "../../fluid-engine-dev_b200/csrc/p2p_ops.cuh"
#include "../../fluid-engine-dev_b200/csrc/diffusion_ops.cuh"

static GridD make_grid(const jetcu_grid* g) {
    return GridD{(int)g->nx, (int)g->ny, (int)g->nz, g->hx, g->hy, g->hz, g->ox, g->oy, g->oz,
                 jc_pow2_inverse(g->hx), jc_pow2_inverse(g->hy), jc_pow2_inverse(g->hz)};
}

// what hyb_extrapolate_one does: `depth` sweeps of k_extrapolate_sweep with ping-pong validity maps
static void extrapolate(double* data, const char* marker, int sx, int sy, int sz, unsigned depth) {
    size_t n = (size_t)sx * sy * sz;
    std::vector<char> v0(marker, marker + n), v1(n, 0);
    for (unsigned it = 0; it < depth; ++it) {
        for (int k = 0; k < sz; ++k)
            for (int j = 0; j < sy; ++j)
                for (int i = 0; i < sx; ++i) extrapolate_cell(data, v0.data(), v1.data(), sx, sy, sz, i, j, k);
        v0.swap(v1);
    }
}

template <int AXIS>
static void markers(const GridD& g, const ColliderD& c, const double* csdf, double* data, char* marker) {
    for (int k = 0; k < g.nz + (AXIS == 2); ++k)
        for (int j = 0; j < g.ny + (AXIS == 1); ++j)
            for (int i = 0; i < g.nx + (AXIS == 0); ++i) constrain_marker_face<AXIS>(g, c, csdf, data, marker, i, j, k);
}

template <int AXIS>
static void project(const GridD& g, const ColliderD& c, const double* csdf, const double* u, const double* v, const double* w,
                    double* temp) {
    for (int k = 0; k < g.nz + (AXIS == 2); ++k)
        for (int j = 0; j < g.ny + (AXIS == 1); ++j)
            for (int i = 0; i < g.nx + (AXIS == 0); ++i) constrain_project_face<AXIS>(g, c, csdf, u, v, w, temp, i, j, k);
}

// ---- G2P and advection: the bodies of k_g2p_pic / k_g2p_flip / k_g2p_apic / k_move over n AoS particles.
// variant: bit 0 = POW2 (exact-reciprocal normalisation; only legal for power-of-two spacings), bit 1 = 64-bit indices.
template <bool POW2, typename I>
static void g2p_pic_t(const GridD& g, const double* u, const double* v, const double* w, uint64_t n, const double* pos, double* vel) {
    for (uint64_t p = 0; p < n; ++p)
        sample_velocity<POW2, I>(g, u, v, w, pos[3 * p], pos[3 * p + 1], pos[3 * p + 2], vel + 3 * p, vel + 3 * p + 1, vel + 3 * p + 2);
}
// as k_move runs it: on PAIR arrays built by k_pack_pairs' body.  variant bit 2: the fused PIC kernel's form (k_pic_move) —
// the grid velocity at the particle becomes its velocity and stands in for the first sample of the first sub-step.
template <bool POW2, typename I>
static void move_t(const GridD& g, const double* u, const double* v, const double* w, double dt, unsigned nsub, int flag,
                   uint64_t n, double* pos, double* vel, bool pic_fused) {
    const size_t nf[3] = {(size_t)(g.nx + 1) * g.ny * g.nz, (size_t)g.nx * (g.ny + 1) * g.nz, (size_t)g.nx * g.ny * (g.nz + 1)};
    const double* in[3] = {u, v, w};
    std::vector<double> pr[3];
    for (int q = 0; q < 3; ++q) {
        pr[q].resize(2 * nf[q]);
        for (size_t i = 0; i < nf[q]; ++i) pack_pair(in[q], pr[q].data(), i, g.nx + (q == 0));
    }
    // even particles through the corner cache (the kernels' default, CornerCache in particle_ops.cuh), odd ones without
    for (uint64_t p = 0; p < n; ++p) {
        CornerCache<I> cache;
        CornerCache<I>* cc = (p & 1) ? nullptr : &cache;
        if (pic_fused) {
            double* vp = vel + 3 * p;
            sample_velocity<POW2, I, true>(g, pr[0].data(), pr[1].data(), pr[2].data(), pos[3 * p], pos[3 * p + 1], pos[3 * p + 2], vp, vp + 1, vp + 2, cc);
            const double first[3] = {vp[0], vp[1], vp[2]};
            move_particle<POW2, I, true>(g, pr[0].data(), pr[1].data(), pr[2].data(), dt, nsub, flag, pos + 3 * p, vp, first, cc);
        } else {
            move_particle<POW2, I, true>(g, pr[0].data(), pr[1].data(), pr[2].data(), dt, nsub, flag, pos + 3 * p, vel + 3 * p, nullptr, cc);
        }
    }
}

// ---- P2G: jetcu_hybrid_transfer_particles_to_grids with k_p2g_tile<APIC, TZ>: the 8 colour passes, the tiles of a pass
// and the "threads" of a tile run one after the other, shared memory is a plain buffer.  Cell planes [kb, ke) carry
// particles (slab mode: the rank's slab + ghost planes), exactly as in the kernel.
template <bool APIC, int AX, int TZ>
static void p2g_tile_component_host(const GridD& g, int I0, int J0, int K0, int kb, int ke, const std::vector<uint32_t>& cell_start,
                                    const AttrPtrs& a, std::vector<double>& ex, std::vector<double>& exw, std::vector<unsigned>& tm,
                                    double* U, double* W, char* M, int* err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    constexpr int WX = P2G_TX + (AX == 0 ? 1 : 2), WY = P2G_TY + (AX == 1 ? 1 : 2), WZ = TZ + (AX == 2 ? 1 : 2);
    for (int tid = 0; tid < NT; ++tid) {
        const int ci = I0 + tid % P2G_TX, cj = J0 + (tid / P2G_TX) % P2G_TY, ck = K0 + tid / (P2G_TX * P2G_TY);
        uint32_t s0 = 0, e0 = 0;
        if (ci < g.nx && cj < g.ny && ck >= kb && ck < ke) {
            const size_t cell = ((size_t)ck * g.ny + cj) * g.nx + ci;
            s0 = cell_start[cell]; e0 = cell_start[cell + 1];
        }
        unsigned touched = 0;
        if (e0 > s0) {
            double sv[P2G_SLOTS], sw[P2G_SLOTS];
            const P2GGlobalSrc<AX> src{&a};
            // both forms of the cell loop (software-pipelined loads / plain), alternating by cell
            touched = (tid & 1) ? p2g_cell_partial<APIC, AX, decltype(src), true>(g, ci, cj, ck, s0, e0, src, sv, sw, err)
                                : p2g_cell_partial<APIC, AX, decltype(src), false>(g, ci, cj, ck, s0, e0, src, sv, sw, err);
            for (int q = 0; q < P2G_SLOTS; ++q) { ex[q * NT + tid] = sv[q]; exw[q * NT + tid] = sw[q]; }
        }
        tm[tid] = touched;
    }
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    const int fx0 = I0 - (AX == 0 ? 0 : 1), fy0 = J0 - (AX == 1 ? 0 : 1), fz0 = K0 - (AX == 2 ? 0 : 1);
    for (int f = 0; f < WX * WY * WZ; ++f) {
        const int lx = f % WX, ly = (f / WX) % WY, lz = f / (WX * WY);
        const int fi = fx0 + lx, fj = fy0 + ly, fk = fz0 + lz;
        if ((unsigned)fi >= (unsigned)sx || (unsigned)fj >= (unsigned)sy || (unsigned)fk >= (unsigned)sz) continue;
        double v, w;
        if (p2g_face_gather<AX, TZ>(ex.data(), tm.data(), lx, ly, lz, &v)) {
            p2g_face_gather<AX, TZ>(exw.data(), tm.data(), lx, ly, lz, &w);
            const size_t idx = ((size_t)fk * sy + fj) * sx + fi;
            U[idx] += v; W[idx] += w; M[idx] = 1;
        }
    }
}

template <bool APIC, int TZ>
static void p2g_tiles(const GridD& g, int kb, int ke, const std::vector<uint32_t>& cell_start, const AttrPtrs& a, double* val[3],
                      double* wsum[3], char* mk[3], int* err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    std::vector<double> ex((size_t)P2G_SLOTS * NT, 0.0), exw((size_t)P2G_SLOTS * NT, 0.0);  // the kernel reuses one buffer
    std::vector<unsigned> tm(NT, 0u);
    const int ntx = (g.nx + P2G_TX - 1) / P2G_TX, nty = (g.ny + P2G_TY - 1) / P2G_TY, tzlo = kb / TZ, tzhi = (ke - 1) / TZ;
    for (int colour = 0; colour < 8; ++colour) {
        const int par_x = colour & 1, par_y = (colour >> 1) & 1, par_z = colour >> 2;
        for (int tz = tzlo + ((tzlo & 1) != par_z ? 1 : 0); tz <= tzhi; tz += 2)
            for (int ty = par_y; ty < nty; ty += 2)
                for (int tx = par_x; tx < ntx; tx += 2) {
                    const int I0 = tx * P2G_TX, J0 = ty * P2G_TY, K0 = tz * TZ;
                    p2g_tile_component_host<APIC, 0, TZ>(g, I0, J0, K0, kb, ke, cell_start, a, ex, exw, tm, val[0], wsum[0], mk[0], err);
                    p2g_tile_component_host<APIC, 1, TZ>(g, I0, J0, K0, kb, ke, cell_start, a, ex, exw, tm, val[1], wsum[1], mk[1], err);
                    p2g_tile_component_host<APIC, 2, TZ>(g, I0, J0, K0, kb, ke, cell_start, a, ex, exw, tm, val[2], wsum[2], mk[2], err);
                }
    }
}

// ---- the staggered-lattice kernel k_p2g_stag<APIC, TZ>: same replay (colour passes, tiles, "threads", plain buffers)
template <bool APIC, int AX, int TZ>
static void p2g_stag_component_host(const GridD& g, int I0, int J0, int K0, int kb, int ke, const std::vector<uint32_t>& sub_start,
                                    const AttrPtrs& a, std::vector<double>& ex, std::vector<unsigned char>& tm, double* U, double* W,
                                    char* M, int* err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    constexpr int WX = P2G_TX + 1, WY = P2G_TY + 1, WZ = TZ + 1;
    const int NI = p2g_stag_count(g.nx, AX == 0), NJ = p2g_stag_count(g.ny, AX == 1), NK = p2g_stag_count(g.nz, AX == 2);
    bool tile_any = false;
    for (int tid = 0; tid < NT; ++tid) {
        const int I = I0 + tid % P2G_TX, J = J0 + (tid / P2G_TX) % P2G_TY, K = K0 + tid / (P2G_TX * P2G_TY);
        bool any = false;
        if (I < NI && J < NJ && K < NK) {
            double sv[8], sw[8];
            const P2GGlobalSrc<AX> src{&a};
            any = p2g_stag_partial<APIC, AX, P2GGlobalSrc<AX>>(g, I, J, K, kb, ke, sub_start.data(), src, sv, sw, err);
            if (any) for (int q = 0; q < 8; ++q) { ex[q * NT + tid] = sv[q]; ex[(8 + q) * NT + tid] = sw[q]; }
        }
        tm[tid] = any ? 1 : 0;
        tile_any |= any;
    }
    if (!tile_any) return;
    const int sx = g.nx + (AX == 0), sy = g.ny + (AX == 1), sz = g.nz + (AX == 2);
    for (int f = 0; f < WX * WY * WZ; ++f) {
        const int lx = f % WX, ly = (f / WX) % WY, lz = f / (WX * WY);
        const int fi = I0 + lx, fj = J0 + ly, fk = K0 + lz;
        if (fi >= sx || fj >= sy || fk >= sz) continue;
        double v, w;
        if (p2g_stag_gather<TZ>(ex.data(), tm.data(), lx, ly, lz, &v)) {
            p2g_stag_gather<TZ>(ex.data() + 8 * NT, tm.data(), lx, ly, lz, &w);
            const size_t idx = ((size_t)fk * sy + fj) * sx + fi;
            U[idx] += v; W[idx] += w; M[idx] = 1;
        }
    }
}

template <bool APIC, int TZ>
static void p2g_stag_tiles(const GridD& g, int kb, int ke, const std::vector<uint32_t>& sub_start, const AttrPtrs& a, double* val[3],
                           double* wsum[3], char* mk[3], int* err) {
    constexpr int NT = P2G_TX * P2G_TY * TZ;
    std::vector<double> ex((size_t)16 * NT, 0.0);
    std::vector<unsigned char> tm(NT, 0);
    const int ntx = (g.nx + P2G_TX - 1) / P2G_TX, nty = (g.ny + P2G_TY - 1) / P2G_TY;
    const int tzlo = std::max(kb - 1, 0) / TZ, tzhi = (ke - 1) / TZ;
    for (int colour = 0; colour < 8; ++colour) {
        const int par_x = colour & 1, par_y = (colour >> 1) & 1, par_z = colour >> 2;
        for (int tz = tzlo + ((tzlo & 1) != par_z ? 1 : 0); tz <= tzhi; tz += 2)
            for (int ty = par_y; ty < nty; ty += 2)
                for (int tx = par_x; tx < ntx; tx += 2) {
                    const int I0 = tx * P2G_TX, J0 = ty * P2G_TY, K0 = tz * TZ;
                    p2g_stag_component_host<APIC, 0, TZ>(g, I0, J0, K0, kb, ke, sub_start, a, ex, tm, val[0], wsum[0], mk[0], err);
                    p2g_stag_component_host<APIC, 1, TZ>(g, I0, J0, K0, kb, ke, sub_start, a, ex, tm, val[1], wsum[1], mk[1], err);
                    p2g_stag_component_host<APIC, 2, TZ>(g, I0, J0, K0, kb, ke, sub_start, a, ex, tm, val[2], wsum[2], mk[2], err);
                }
    }
}

extern "C" {

// PIC / FLIP (apic = 0) or APIC (apic = 1) transferFromParticlesToGrids through the tile kernel's source with tiles of tz
// (4 or 8) cell planes.  slabs > 1 replays the z-slab mode: every "rank" sees only the particles of its slab + 2 ghost
// planes and finalises the face planes it owns — the result must not depend on slabs.  AoS inputs in the caller's order.
// tz = 0 selects the staggered-lattice kernel (k_p2g_stag, the product's default), tz = 4 / 8 the 18-slot window kernel.
int hc_p2g(const jetcu_grid* grid, int apic, int tz, int slabs, uint64_t n, const double* pos, const double* vel, const double* cx,
           const double* cy, const double* cz, double* u, double* v, double* w, char* um, char* vm, char* wm, int* err) {
    GridD g = make_grid(grid);
    const size_t nc = (size_t)g.nx * g.ny * g.nz;
    const size_t pl[3] = {(size_t)(g.nx + 1) * g.ny, (size_t)g.nx * (g.ny + 1), (size_t)g.nx * g.ny};
    const size_t nf[3] = {pl[0] * g.nz, pl[1] * g.nz, pl[2] * (g.nz + 1)};
    if (slabs < 1 || g.nz % slabs != 0 || (tz != 0 && tz != 4 && tz != 8)) return 1;
    // bin sort: by (owning cell, octant), ascending original index inside a bin (k_cell_count .. k_sort_cells .. k_gather_attrs)
    std::vector<uint32_t> cell(n), order(n), bin(n);
    std::vector<int> plane(n);
    for (uint64_t p = 0; p < n; ++p) {
        const int ci = cell_coord(pos[3 * p], g.ox, 1.0 / g.hx, g.nx), cj = cell_coord(pos[3 * p + 1], g.oy, 1.0 / g.hy, g.ny),
                  ck = cell_coord(pos[3 * p + 2], g.oz, 1.0 / g.hz, g.nz);
        cell[p] = (uint32_t)(((size_t)ck * g.ny + cj) * g.nx + ci);
        bin[p] = 8 * cell[p] + (uint32_t)p2g_octant(g, pos[3 * p], pos[3 * p + 1], pos[3 * p + 2], ci, cj, ck);
        plane[p] = ck;
    }
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return bin[x] < bin[y]; });
    double* out[3] = {u, v, w};
    char* outm[3] = {um, vm, wm};
    *err = 0;
    const int T = g.nz / slabs;
    for (int rank = 0; rank < slabs; ++rank) {
        const int z0 = rank * T, z1 = z0 + T;
        const int kb = slabs > 1 ? std::max(z0 - 2, 0) : 0, ke = slabs > 1 ? std::min(z1 + 2, g.nz) : g.nz;
        // this rank's sorted particle array: the particles of the planes [kb, ke) only
        std::vector<uint32_t> mine;
        for (uint64_t s = 0; s < n; ++s) if (plane[order[s]] >= kb && plane[order[s]] < ke) mine.push_back(order[s]);
        const size_t m = mine.size();
        std::vector<uint32_t> cell_start(nc + 1, 0);
        for (size_t s = 0; s < m; ++s) cell_start[cell[mine[s]] + 1]++;
        for (size_t c = 0; c < nc; ++c) cell_start[c + 1] += cell_start[c];
        std::vector<uint32_t> sub_start(8 * nc + 1, 0);
        for (size_t s = 0; s < m; ++s) sub_start[bin[mine[s]] + 1]++;
        for (size_t c = 0; c < 8 * nc; ++c) sub_start[c + 1] += sub_start[c];
        std::vector<std::vector<double>> soa(NATTR, std::vector<double>(m ? m : 1, 0.0));
        for (size_t s = 0; s < m; ++s) {
            const uint32_t p = mine[s];
            for (int q = 0; q < 3; ++q) {
                soa[q][s] = pos[3 * p + q]; soa[3 + q][s] = vel[3 * p + q];
                if (apic) { soa[6 + q][s] = cx[3 * p + q]; soa[9 + q][s] = cy[3 * p + q]; soa[12 + q][s] = cz[3 * p + q]; }
            }
        }
        AttrPtrs a;
        for (int q = 0; q < NATTR; ++q) a.p[q] = soa[q].data();
        std::vector<double> valv[3], wsv[3];
        std::vector<char> mkv[3];
        double *val[3], *ws[3];
        char* mk[3];
        for (int q = 0; q < 3; ++q) {
            valv[q].assign(nf[q], 0.0); wsv[q].assign(nf[q], 0.0); mkv[q].assign(nf[q], 0);
            val[q] = valv[q].data(); ws[q] = wsv[q].data(); mk[q] = mkv[q].data();
        }
        if (m > 0 && tz == 0) {
            if (apic) p2g_stag_tiles<true, 4>(g, kb, ke, sub_start, a, val, ws, mk, err);
            else p2g_stag_tiles<false, 4>(g, kb, ke, sub_start, a, val, ws, mk, err);
        } else if (m > 0) {
            if (apic) { if (tz == 4) p2g_tiles<true, 4>(g, kb, ke, cell_start, a, val, ws, mk, err); else p2g_tiles<true, 8>(g, kb, ke, cell_start, a, val, ws, mk, err); }
            else { if (tz == 4) p2g_tiles<false, 4>(g, kb, ke, cell_start, a, val, ws, mk, err); else p2g_tiles<false, 8>(g, kb, ke, cell_start, a, val, ws, mk, err); }
        }
        // the planes this rank owns (the last rank also owns the top w plane)
        for (int q = 0; q < 3; ++q) {
            const size_t f0 = (size_t)z0 * pl[q], f1 = ((size_t)z1 + ((q == 2 && rank == slabs - 1) ? 1 : 0)) * pl[q];
            for (size_t f = f0; f < f1; ++f) {
                p2g_normalise_face(val[q], ws[q], nullptr, f);
                out[q][f] = val[q][f]; outm[q][f] = mk[q][f];
            }
        }
    }
    return 0;
}

// k_collider_sdf over the whole grid
int hc_collider_sdf(const jetcu_grid* grid, const jetcu_collider_shape* shapes, uint32_t ns, int as_set, const double* lin, const double* ang,
                    double friction, double* sdf) {
    GridD g = make_grid(grid);
    ColliderD c = cd_make_collider(shapes, ns, as_set, lin, ang, friction);
    for (int k = 0; k < g.nz; ++k)
        for (int j = 0; j < g.ny; ++j)
            for (int i = 0; i < g.nx; ++i) sdf[((size_t)k * g.ny + j) * g.nx + i] = collider_sdf_cell(g, c, i, j, k);
    return 0;
}

// hyb_zero_walls() of hybrid.cu == GridFluidSolver3::applyBoundaryCondition: hyb_constrain_collider, then the walls
int hc_apply_boundary_condition(const jetcu_grid* grid, const jetcu_collider_shape* shapes, uint32_t ns, int as_set, const double* lin,
                                const double* ang, double friction, double max_cfl, int wall_flag, double* u, double* v,
                                double* w) {
    GridD g = make_grid(grid);
    const size_t nu = (size_t)(g.nx + 1) * g.ny * g.nz, nv = (size_t)g.nx * (g.ny + 1) * g.nz, nw = (size_t)g.nx * g.ny * (g.nz + 1);
    if (ns != 0) {
        ColliderD c = cd_make_collider(shapes, ns, as_set, lin, ang, friction);
        std::vector<double> csdf((size_t)g.nx * g.ny * g.nz), tu(nu), tv(nv), tw(nw);
        hc_collider_sdf(grid, shapes, ns, as_set, lin, ang, friction, csdf.data());
        std::vector<char> mu(nu), mv(nv), mw(nw);
        const unsigned depth = (unsigned)ceil(max_cfl);
        markers<0>(g, c, csdf.data(), u, mu.data());
        markers<1>(g, c, csdf.data(), v, mv.data());
        markers<2>(g, c, csdf.data(), w, mw.data());
        extrapolate(u, mu.data(), g.nx + 1, g.ny, g.nz, depth);
        extrapolate(v, mv.data(), g.nx, g.ny + 1, g.nz, depth);
        extrapolate(w, mw.data(), g.nx, g.ny, g.nz + 1, depth);
        project<0>(g, c, csdf.data(), u, v, w, tu.data());
        project<1>(g, c, csdf.data(), u, v, w, tv.data());
        project<2>(g, c, csdf.data(), u, v, w, tw.data());
        memcpy(u, tu.data(), nu * 8); memcpy(v, tv.data(), nv * 8); memcpy(w, tw.data(), nw * 8);
    }
    size_t m = (size_t)g.ny * g.nz;
    if ((size_t)g.nx * g.nz > m) m = (size_t)g.nx * g.nz;
    if ((size_t)g.nx * g.ny > m) m = (size_t)g.nx * g.ny;
    for (size_t t = 0; t < m; ++t) zero_walls_elem(g, wall_flag, u, v, w, t);
    return 0;
}

// boundary velocity samples handed to the pressure solver (k_boundary_velocity<AXIS>)
int hc_boundary_velocity(const jetcu_grid* grid, const jetcu_collider_shape* shapes, uint32_t ns, int as_set, const double* lin, const double* ang,
                         double friction, double* bu, double* bv, double* bw) {
    GridD g = make_grid(grid);
    ColliderD c = cd_make_collider(shapes, ns, as_set, lin, ang, friction);
    for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx + 1; ++i) boundary_velocity_face<0>(g, c, bu, i, j, k);
    for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny + 1; ++j) for (int i = 0; i < g.nx; ++i) boundary_velocity_face<1>(g, c, bv, i, j, k);
    for (int k = 0; k < g.nz + 1; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx; ++i) boundary_velocity_face<2>(g, c, bw, i, j, k);
    return 0;
}

// the extrapolateIntoCollider tail of jetcu_hybrid_build_signed_distance_field: k_outside_collider_marker + sweeps
int hc_extrapolate_into_collider(const jetcu_grid* grid, const jetcu_collider_shape* shapes, uint32_t ns, int as_set, const double* lin,
                                 const double* ang, double friction, double max_cfl, double* sdf) {
    GridD g = make_grid(grid);
    std::vector<double> csdf((size_t)g.nx * g.ny * g.nz);
    hc_collider_sdf(grid, shapes, ns, as_set, lin, ang, friction, csdf.data());
    std::vector<char> marker(csdf.size());
    for (int k = 0; k < g.nz; ++k)
        for (int j = 0; j < g.ny; ++j)
            for (int i = 0; i < g.nx; ++i) marker[((size_t)k * g.ny + j) * g.nx + i] = outside_collider_cell(g, csdf.data(), i, j, k);
    extrapolate(sdf, marker.data(), g.nx, g.ny, g.nz, (unsigned)ceil(max_cfl));
    return 0;
}

// jetcu_hybrid_emit_shapes: cd_emit_region + the lattice loop of hyb_emit_lattice with k_emit_flags_shapes' acceptance test.
// Call with pos == NULL to count; returns the number of accepted points through *n_out.
int hc_emit_shapes(const jetcu_grid* g, const jetcu_collider_shape* shapes, uint32_t ns, int as_set, int bcc, double spacing, double* pos,
                   uint64_t* n_out) {
    const ColliderD e = cd_make_collider(shapes, ns, as_set, nullptr, nullptr, 0.0);
    double lo[3] = {g->ox, g->oy, g->oz};
    double hi[3] = {g->ox + g->hx * (double)g->nx, g->oy + g->hy * (double)g->ny, g->oz + g->hz * (double)g->nz};
    cd_emit_region(e, lo, hi);
    LatticeD L;
    const size_t total = cd_lattice_make(lo, hi, spacing, bcc, &L);
    uint64_t o = 0;
    for (size_t t = 0; t < total; ++t) {
        double p[3];
        lattice_point_l(L, t, p);
        if (!cd_emit_accept(e, p)) continue;
        if (pos) { pos[3 * o] = p[0]; pos[3 * o + 1] = p[1]; pos[3 * o + 2] = p[2]; }
        ++o;
    }
    *n_out = o;
    return 0;
}

// k_resolve_collision over n AoS particles
int hc_resolve_collision(const jetcu_collider_shape* shapes, uint32_t ns, int as_set, const double* lin, const double* ang, double friction,
                         uint64_t n, double* pos, double* vel) {
    ColliderD c = cd_make_collider(shapes, ns, as_set, lin, ang, friction);
    for (uint64_t p = 0; p < n; ++p) cd_resolve_collision(c, pos + 3 * p, vel + 3 * p);
    return 0;
}

// ---- G2P and advection entry points (templates above)
int hc_g2p_pic(const jetcu_grid* grid, const double* u, const double* v, const double* w, int variant, uint64_t n,
               const double* pos, double* vel) {
    GridD g = make_grid(grid);
    if ((variant & 1) && !(g.ihx != 0.0 && g.ihy != 0.0 && g.ihz != 0.0)) return 1;
    switch (variant & 3) {
        case 0: g2p_pic_t<false, int>(g, u, v, w, n, pos, vel); break;
        case 1: g2p_pic_t<true, int>(g, u, v, w, n, pos, vel); break;
        case 2: g2p_pic_t<false, long long>(g, u, v, w, n, pos, vel); break;
        default: g2p_pic_t<true, long long>(g, u, v, w, n, pos, vel); break;
    }
    return 0;
}

int hc_g2p_flip(const jetcu_grid* grid, const double* u, const double* v, const double* w, const float* ud, const float* vd,
                const float* wd, double blend, uint64_t n, const double* pos, double* vel) {
    GridD g = make_grid(grid);
    for (uint64_t p = 0; p < n; ++p) g2p_flip_particle(g, u, v, w, ud, vd, wd, blend, pos + 3 * p, vel + 3 * p);
    return 0;
}

int hc_g2p_apic(const jetcu_grid* grid, const double* u, const double* v, const double* w, uint64_t n, const double* pos,
                double* vel, double* cx, double* cy, double* cz) {
    GridD g = make_grid(grid);
    const bool pow2 = g.ihx != 0.0 && g.ihy != 0.0 && g.ihz != 0.0;
    for (uint64_t p = 0; p < n; ++p) {
        const double px = pos[3 * p], py = pos[3 * p + 1], pz = pos[3 * p + 2];
        if (pow2) sample_velocity<true, int>(g, u, v, w, px, py, pz, vel + 3 * p, vel + 3 * p + 1, vel + 3 * p + 2);
        else sample_velocity<false, int>(g, u, v, w, px, py, pz, vel + 3 * p, vel + 3 * p + 1, vel + 3 * p + 2);
        apic_affine<0>(g, u, px, py, pz, cx + 3 * p, cx + 3 * p + 1, cx + 3 * p + 2);
        apic_affine<1>(g, v, px, py, pz, cy + 3 * p, cy + 3 * p + 1, cy + 3 * p + 2);
        apic_affine<2>(g, w, px, py, pz, cz + 3 * p, cz + 3 * p + 1, cz + 3 * p + 2);
    }
    return 0;
}

int hc_move_particles(const jetcu_grid* grid, const double* u, const double* v, const double* w, double dt_frame, double max_cfl,
                      int flag, int variant, uint64_t n, double* pos, double* vel) {
    GridD g = make_grid(grid);
    if ((variant & 1) && !(g.ihx != 0.0 && g.ihy != 0.0 && g.ihz != 0.0)) return 1;
    const unsigned nsub = (unsigned)fmax(max_cfl, 1.0);   // jetcu_hybrid_move_particles
    const double dt = dt_frame / nsub;
    switch (variant & 3) {
        case 0: move_t<false, int>(g, u, v, w, dt, nsub, flag, n, pos, vel, (variant & 4) != 0); break;
        case 1: move_t<true, int>(g, u, v, w, dt, nsub, flag, n, pos, vel, (variant & 4) != 0); break;
        case 2: move_t<false, long long>(g, u, v, w, dt, nsub, flag, n, pos, vel, (variant & 4) != 0); break;
        default: move_t<true, long long>(g, u, v, w, dt, nsub, flag, n, pos, vel, (variant & 4) != 0); break;
    }
    return 0;
}

// ---- GridSmokeSolver3 stages: the bodies of smoke.cu's kernels behind plain loops (host code of smoke.cu replayed)
static void smoke_walls(const GridD& g, int flag, double* u, double* v, double* w) {
    size_t m = (size_t)g.ny * g.nz;
    if ((size_t)g.nx * g.nz > m) m = (size_t)g.nx * g.nz;
    if ((size_t)g.nx * g.ny > m) m = (size_t)g.nx * g.ny;
    for (size_t t = 0; t < m; ++t) zero_walls_elem(g, flag, u, v, w, t);
}

// jetcu_smoke_compute_external_forces with the ambient temperature SUM supplied by the caller (the GPU reduces it with
// a tree, the reference serially; the per-face arithmetic is what is checked here)
int hc_smoke_buoyancy(const jetcu_grid* grid, const double* den, const double* temp, double fd, double ft, double t_sum,
                      const double* gravity, double dt, int wall_flag, double* u, double* v, double* w) {
    GridD g = make_grid(grid);
    double up[3] = {0, 1, 0};
    if (gravity[0] * gravity[0] + gravity[1] * gravity[1] + gravity[2] * gravity[2] > DBL_EPSILON) {
        const double l = sqrt(gravity[0] * gravity[0] + gravity[1] * gravity[1] + gravity[2] * gravity[2]);
        up[0] = -(gravity[0] / l); up[1] = -(gravity[1] / l); up[2] = -(gravity[2] / l);
    }
    if (!(fabs(fd) > DBL_EPSILON || fabs(ft) > DBL_EPSILON)) return 0;
    const double t_amb = t_sum / (double)((size_t)g.nx * g.ny * g.nz);
    if (fabs(up[0]) > DBL_EPSILON)
        for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx + 1; ++i)
            sm_buoyancy_face<0>(g, den, temp, fd, ft, t_amb, dt, up[0], u, i, j, k);
    if (fabs(up[1]) > DBL_EPSILON)
        for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny + 1; ++j) for (int i = 0; i < g.nx; ++i)
            sm_buoyancy_face<1>(g, den, temp, fd, ft, t_amb, dt, up[1], v, i, j, k);
    if (fabs(up[2]) > DBL_EPSILON)
        for (int k = 0; k < g.nz + 1; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx; ++i)
            sm_buoyancy_face<2>(g, den, temp, fd, ft, t_amb, dt, up[2], w, i, j, k);
    smoke_walls(g, wall_flag, u, v, w);
    return 0;
}

// jetcu_smoke_compute_advection
// k_smoke_emit_step over the whole grid (jetcu_smoke_emit_step applies it to density and temperature)
int hc_smoke_emit_step(const jetcu_grid* grid, const jetcu_collider_shape* shapes, uint32_t ns, int as_set, double lo, double hi,
                       double* field) {
    GridD g = make_grid(grid);
    const ColliderD e = cd_make_collider(shapes, ns, as_set, nullptr, nullptr, 0.0);
    const double width = cd_min(cd_min(g.hx, g.hy), g.hz);
    for (int k = 0; k < g.nz; ++k)
        for (int j = 0; j < g.ny; ++j)
            for (int i = 0; i < g.nx; ++i) sm_emit_step_cell(g, e, width, lo, hi, field, i, j, k);
    return 0;
}

// jetcu_smoke_cfl: k_smoke_cfl_partial's body over all cells
int hc_smoke_cfl(const jetcu_grid* grid, const double* u, const double* v, const double* w, const double* gravity, double dt, double* cfl) {
    GridD g = make_grid(grid);
    double m = 0.0;
    const size_t nc = (size_t)g.nx * g.ny * g.nz;
    for (size_t c = 0; c < nc; ++c) m = fmax(m, sm_cfl_cell(g, u, v, w, dt * gravity[0], dt * gravity[1], dt * gravity[2], c));
    *cfl = m * dt / fmin(fmin(g.hx, g.hy), g.hz);
    return 0;
}

// jetcu_smoke_compute_advection; ns > 0: with a rigid collider (boundary handling in the back-trace, cells inside the collider not
// advected, extrapolateIntoCollider of the scalars, the collider branch of the boundary condition afterwards)
int hc_smoke_advect_collider(const jetcu_grid* grid, double dt, int wall_flag, int linear_i, const jetcu_collider_shape* shapes,
                             uint32_t ns, int as_set, const double* lin, const double* ang, double friction, double max_cfl, double* den,
                             double* temp, double* u, double* v, double* w) {
    const bool linear = linear_i != 0;
    GridD g = make_grid(grid);
    const size_t nc = (size_t)g.nx * g.ny * g.nz;
    const size_t nu = (size_t)(g.nx + 1) * g.ny * g.nz, nv = (size_t)g.nx * (g.ny + 1) * g.nz, nw = (size_t)g.nx * g.ny * (g.nz + 1);
    const double hmin = fmin(fmin(g.hx, g.hy), g.hz);
    std::vector<double> sdf_store;
    const double* csdf = nullptr;
    if (ns > 0) {
        sdf_store.resize(nc);
        hc_collider_sdf(grid, shapes, ns, as_set, lin, ang, friction, sdf_store.data());
        csdf = sdf_store.data();
    }
    double* fields[2] = {den, temp};
    for (double* f : fields) {
        std::vector<double> in(f, f + nc);
        for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx; ++i)
            f[((size_t)k * g.ny + j) * g.nx + i] = sm_advect_cell(g, in.data(), u, v, w, dt, hmin, i, j, k, linear, csdf);
        if (csdf) {
            std::vector<char> marker(nc);
            for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx; ++i)
                marker[((size_t)k * g.ny + j) * g.nx + i] = outside_collider_cell(g, csdf, i, j, k);
            extrapolate(f, marker.data(), g.nx, g.ny, g.nz, (unsigned)ceil(max_cfl));
        }
    }
    std::vector<double> u0(u, u + nu), v0(v, v + nv), w0(w, w + nw);
    for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx + 1; ++i)
        u[((size_t)k * g.ny + j) * (g.nx + 1) + i] = sm_advect_face<0>(g, u0.data(), v0.data(), w0.data(), dt, hmin, i, j, k, linear, csdf);
    for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny + 1; ++j) for (int i = 0; i < g.nx; ++i)
        v[((size_t)k * (g.ny + 1) + j) * g.nx + i] = sm_advect_face<1>(g, u0.data(), v0.data(), w0.data(), dt, hmin, i, j, k, linear, csdf);
    for (int k = 0; k < g.nz + 1; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx; ++i)
        w[((size_t)k * g.ny + j) * g.nx + i] = sm_advect_face<2>(g, u0.data(), v0.data(), w0.data(), dt, hmin, i, j, k, linear, csdf);
    if (ns > 0) return hc_apply_boundary_condition(grid, shapes, ns, as_set, lin, ang, friction, max_cfl, wall_flag, u, v, w);
    smoke_walls(g, wall_flag, u, v, w);
    return 0;
}

int hc_smoke_advect(const jetcu_grid* grid, double dt, int wall_flag, int linear_i, double* den, double* temp, double* u, double* v, double* w) {
    return hc_smoke_advect_collider(grid, dt, wall_flag, linear_i, nullptr, 0, 0, nullptr, nullptr, 0.0, 5.0, den, temp, u, v, w);
}

// ---- peer-memory halo exchange (p2p_ops.cuh): a chain of `world` ranks in host memory, each with its full-size globally indexed
// array (planes of `plane_bytes`, `T` owned planes per rank) and its mailbox; `exchanges` rounds of push (all ranks, in the order
// given by `reverse`) then pull, with the very addressing comm.cu's kernels use.  Owned planes carry a (rank, plane, round) pattern,
// everything else is poisoned before each round.  Returns the number of halo bytes that do not hold the owner's pattern (0 = pass),
// or -1 if a pull would have had to wait (message numbers out of step).
long long hc_p2p_chain(int world, int T, uint64_t plane_bytes, int depth, int exchanges, uint64_t slot_bytes, int reverse) {
    const int nz = world * T;
    std::vector<std::vector<unsigned char>> arr(world, std::vector<unsigned char>((size_t)nz * plane_bytes));
    std::vector<std::vector<unsigned char>> box(world, std::vector<unsigned char>(P2P_HEADER + 4 * slot_bytes, 0));
    auto pattern = [&](int plane, int round, size_t byte) { return (unsigned char)(plane * 7 + round * 31 + (byte % 13) + 1); };
    long long bad = 0;
    for (int round = 0; round < exchanges; ++round) {
        std::vector<P2PExchange> ex(world);
        for (int r = 0; r < world; ++r) {
            std::fill(arr[r].begin(), arr[r].end(), (unsigned char)0xEE);
            for (int k = r * T; k < (r + 1) * T; ++k)
                for (size_t b = 0; b < plane_bytes; ++b) arr[r][(size_t)k * plane_bytes + b] = pattern(k, round, b);
            const bool down = r > 0, up = r + 1 < world;
            ex[r] = p2p_exchange(reinterpret_cast<char*>(box[r].data()), down ? reinterpret_cast<char*>(box[r - 1].data()) : nullptr,
                                 up ? reinterpret_cast<char*>(box[r + 1].data()) : nullptr, slot_bytes, reinterpret_cast<char*>(arr[r].data()),
                                 plane_bytes, r * T, (r + 1) * T, depth, down, up);
        }
        for (int q = 0; q < world; ++q) {   // push (k_p2p_push)
            const int r = reverse ? world - 1 - q : q;
            const P2PExchange& e = ex[r];
            const unsigned long long seq = e.me->seq + 1;
            const size_t off = p2p_slot_offset(e, seq);
            for (size_t i = 0; i < e.n16; ++i) {
                if (e.src_dn) e.box_dn[off + i] = e.src_dn[i];
                if (e.src_up) e.box_up[off + i] = e.src_up[i];
            }
            if (e.flag_dn) *e.flag_dn = seq;
            if (e.flag_up) *e.flag_up = seq;
        }
        for (int r = 0; r < world; ++r) {   // pull (k_p2p_pull)
            const P2PExchange& e = ex[r];
            const unsigned long long seq = e.me->seq + 1;
            if (!p2p_arrived(e, e.me->data[0], e.me->data[1], seq)) return -1;
            const size_t off = p2p_slot_offset(e, seq);
            for (size_t i = 0; i < e.n16; ++i) {
                if (e.dst_dn) e.dst_dn[i] = e.in_dn[off + i];
                if (e.dst_up) e.dst_up[i] = e.in_up[off + i];
            }
            e.me->seq = seq;
        }
        for (int r = 0; r < world; ++r) {   // every halo plane must hold its owner's pattern of this round, nothing else changed
            for (int k = 0; k < nz; ++k) {
                const bool owned = k >= r * T && k < (r + 1) * T;
                const bool halo = (r > 0 && k >= r * T - depth && k < r * T) || (r + 1 < world && k >= (r + 1) * T && k < (r + 1) * T + depth);
                for (size_t b = 0; b < plane_bytes; ++b) {
                    const unsigned char got = arr[r][(size_t)k * plane_bytes + b];
                    const unsigned char want = (owned || halo) ? pattern(k, round, b) : (unsigned char)0xEE;
                    if (got != want) ++bad;
                }
            }
        }
    }
    return bad;
}

// ---- linear-system kernels (fdm_ops.cuh) and the pressure system (pressure_ops.cuh)
}  // extern "C" (templates and state below)

namespace {
struct HostLevel {
    int nx, ny, nz;
    std::vector<float> sdf, uw, vw, ww, cr, cu, cf;
    std::vector<double> center;
    std::vector<uint8_t> code;
    Mat m;
};
struct HostSystem {
    jetcu_grid grid{};
    std::vector<HostLevel> lv;
    std::vector<double> b0;
} g_sys;

struct SoAMat {
    std::vector<double> c, r, u, f;
    Mat m;
    SoAMat(const jetcu_matrix_row* A, size_t n) : c(n), r(n), u(n), f(n) {
        for (size_t i = 0; i < n; ++i) { c[i] = A[i].center; r[i] = A[i].right; u[i] = A[i].up; f[i] = A[i].front; }
        m.center = c.data(); m.right = r.data(); m.up = u.data(); m.front = f.data();
    }
};

template <bool P>
void relax_sweeps(const Mat& A, const double* b, double* x, double w, unsigned iterations, int nx, int ny, int nz, bool code) {
    for (unsigned it = 0; it < iterations; ++it)
        for (int colour = 0; colour < 2; ++colour)   // launch_colour<P>: red = (i+j+k) even, then black
            for (int k = 0; k < nz; ++k)
                for (int j = 0; j < ny; ++j)
                    for (int i = (j + k + colour) & 1; i < nx; i += 2) {
                        if (code) relax_cell_code(A, b, x, w, i, j, k, nx, ny, nz);
                        else relax_cell<P>(A, b, x, w, i, j, k, nx, ny, nz);
                    }
}

// launch_relax_lex of fdm.cu (the natural Gauss-Seidel ordering, FdmGaussSeidelSolver3::relax) replayed with its loop structure.
// Levels of at most 8192 cells: k_relax_lex_small's steps (cell hyperplane h of sweep s in step h + 2 s); larger ones: the
// launches of k_relax_lex_tiles (tile hyperplane t - 2 s for every sweep s in flight, 32 at most), inside a tile the cell
// hyperplanes in order.  CTAs of one launch / threads of one step run concurrently on the device, so the replay visits them
// in ascending (order 0) or descending (order 1) order: both must give the serial loop's bits.
template <bool P>
void relax_lex_schedule(const Mat& A, const double* b, double* x, double w, unsigned iterations, int nx, int ny, int nz, bool code,
                        int order) {
    auto cell = [&](int i, int j, int k) {
        if (code) relax_cell_code(A, b, x, w, i, j, k, nx, ny, nz);
        else relax_cell<P>(A, b, x, w, i, j, k, nx, ny, nz);
    };
    const long long n = (long long)nx * ny * nz;
    if (n <= 8192) {
        const int steps = (nx + ny + nz - 2) + 2 * ((int)iterations - 1);
        for (int step = 0; step < steps; ++step)
            for (long long q = 0; q < n; ++q) {
                const long long c = order ? n - 1 - q : q;
                const int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long long)nx * ny));
                const int lag = step - (i + j + k);
                if (lag < 0 || (lag & 1) || lag / 2 >= (int)iterations) continue;
                relax_cell<P>(A, b, x, w, i, j, k, nx, ny, nz);   // the small kernel never uses the codes
            }
        return;
    }
    const int T = 8, ntx = (nx + T - 1) / T, nty = (ny + T - 1) / T, ntz = (nz + T - 1) / T, last = ntx + nty + ntz - 3;
    for (unsigned done = 0; done < iterations;) {
        const int ns = (int)((iterations - done < 32u) ? iterations - done : 32u);
        for (int t = 0; t <= last + 2 * (ns - 1); ++t) {
            const int s_lo = t - last > 0 ? (t - last + 1) / 2 : 0, s_hi = t / 2 < ns - 1 ? t / 2 : ns - 1;
            const int nsl = s_hi - s_lo + 1;
            if (nsl <= 0) continue;
            const long long ctas = (long long)nsl * nty * ntx;
            // order >= 2: a pseudo-random permutation of the launch's CTAs (q -> (a q + b) mod ctas with gcd(a, ctas) = 1)
            long long pa = 1, pb = 0;
            if (order >= 2) {
                pa = 2654435761ll % ctas; if (pa == 0) pa = 1;
                auto gcd = [](long long u, long long v) { while (v) { const long long r = u % v; u = v; v = r; } return u; };
                while (gcd(pa, ctas) != 1) ++pa;
                pb = (long long)(40503u * (unsigned)(t + 7 * order)) % ctas;
            }
            for (long long q = 0; q < ctas; ++q) {
                const long long cta = order >= 2 ? (pa * q + pb) % ctas : (order ? ctas - 1 - q : q);
                const int ti = (int)(cta % ntx), tj = (int)((cta / ntx) % nty), bz = (int)(cta / ((long long)ntx * nty));
                const int tk = t - 2 * (s_lo + bz) - ti - tj;
                if (tk < 0 || tk >= ntz) continue;
                for (int step = 0; step <= 3 * T - 3; ++step)
                    for (int th = 0; th < T * T * T; ++th) {
                        const int cx = th % T, cy = (th / T) % T, cz = th / (T * T);
                        const int i = ti * T + cx, j = tj * T + cy, k = tk * T + cz;
                        if (i >= nx || j >= ny || k >= nz || cx + cy + cz != step) continue;
                        // the tile kernel's arithmetic: coefficients gathered first, absent neighbours as 0.0 (fdm_ops.cuh)
                        if (code) relax_cell_split<P, true>(A, b, x, w, i, j, k, nx, ny, nz);
                        else relax_cell_split<P, false>(A, b, x, w, i, j, k, nx, ny, nz);
                    }
            }
        }
        done += (unsigned)ns;
    }
}
}  // namespace

extern "C" {

int hc_mvm(int nx, int ny, int nz, const jetcu_matrix_row* A, const double* v, double* out) {
    SoAMat M(A, (size_t)nx * ny * nz);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) mvm_cell<false>(M.m, v, out, i, j, k, nx, ny, nz);
    return 0;
}
int hc_residual(int nx, int ny, int nz, const jetcu_matrix_row* A, const double* x, const double* b, double* out) {
    SoAMat M(A, (size_t)nx * ny * nz);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) residual_cell<false>(M.m, x, b, out, i, j, k, nx, ny, nz);
    return 0;
}
int hc_relax(int nx, int ny, int nz, const jetcu_matrix_row* A, const double* b, double sor, unsigned iterations, double* x) {
    SoAMat M(A, (size_t)nx * ny * nz);
    relax_sweeps<false>(M.m, b, x, sor, iterations, nx, ny, nz, false);
    return 0;
}
int hc_relax_lex(int nx, int ny, int nz, const jetcu_matrix_row* A, const double* b, double sor, unsigned iterations, int order, double* x) {
    SoAMat M(A, (size_t)nx * ny * nz);
    relax_lex_schedule<false>(M.m, b, x, sor, iterations, nx, ny, nz, false, order);
    return 0;
}
// k_ic_tiles<false, OP> of iccg.cu in its launch order: tile hyperplanes one after the other (backward: descending), the
// tiles of a hyperplane in any order (here: ascending), inside a tile the cell hyperplanes one after the other.  FdmIccgSolver3's
// Preconditioner::build (d) and ::solve (y forward, x backward).
int hc_iccg_precondition(int nx, int ny, int nz, const jetcu_matrix_row* A, const double* b, double* d, double* y, double* x) {
    SoAMat M(A, (size_t)nx * ny * nz);
    const int T = 8, ntx = (nx + T - 1) / T, nty = (ny + T - 1) / T, ntz = (nz + T - 1) / T, last = ntx + nty + ntz - 3;
    for (int op = 0; op < 3; ++op)
        for (int s = 0; s <= last; ++s) {
            const int S = op == 2 ? last - s : s;
            for (int tj = 0; tj < nty; ++tj)
                for (int ti = 0; ti < ntx; ++ti) {
                    const int tk = S - ti - tj;
                    if (tk < 0 || tk >= ntz) continue;
                    for (int step = 0; step <= 3 * T - 3; ++step)
                        for (int t = 0; t < T * T * T; ++t) {
                            const int cx = t % T, cy = (t / T) % T, cz = t / (T * T);
                            const int i = ti * T + cx, j = tj * T + cy, k = tk * T + cz;
                            const int mine = op == 2 ? (3 * T - 3) - (cx + cy + cz) : cx + cy + cz;
                            if (i >= nx || j >= ny || k >= nz || mine != step) continue;
                            const size_t c = ((size_t)k * ny + j) * nx + i;
                            if (op == 0) d[c] = ic_build_cell<false>(M.m, d, i, j, k, nx, ny);
                            else if (op == 1) y[c] = ic_forward_cell<false>(M.m, b, y, d, i, j, k, nx, ny);
                            else x[c] = ic_backward_cell<false>(M.m, y, x, d, i, j, k, nx, ny, nz);
                        }
                }
        }
    return 0;
}

int hc_restrict(int nx, int ny, int nz, const double* finer, double* coarser) {   // (nx, ny, nz) = finer size
    const int cx = nx / 2, cy = ny / 2, cz = nz / 2;
    for (int k = 0; k < cz; ++k) for (int j = 0; j < cy; ++j) for (int i = 0; i < cx; ++i) restrict_cell(finer, coarser, i, j, k, cx, cy, cz);
    return 0;
}
int hc_correct(int nx, int ny, int nz, const double* coarser, double* finer) {
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) correct_cell(coarser, finer, i, j, k, nx, ny, nz);
    return 0;
}

// pressure_solve_device's system build (buildWeights, restrict, buildSystem) with the packed matrix, on the host
int hc_pressure_build(const jetcu_grid* grid, uint32_t max_levels, const double* fluid_sdf, const double* collider_sdf,
                      const double* u, const double* v, const double* w, const double* bu, const double* bv, const double* bw) {
    g_sys = HostSystem();
    g_sys.grid = *grid;
    GridP g{(int)grid->nx, (int)grid->ny, (int)grid->nz, grid->hx, grid->hy, grid->hz, grid->ox, grid->oy, grid->oz};
    uint32_t levels = 1;
    {
        uint64_t nx = grid->nx, ny = grid->ny, nz = grid->nz;
        for (; levels < max_levels; ++levels) {
            if (nx % 2 == 0 && ny % 2 == 0 && nz % 2 == 0) { nx >>= 1; ny >>= 1; nz >>= 1; } else break;
        }
    }
    g_sys.lv.resize(levels);
    for (uint32_t l = 0; l < levels; ++l) {
        HostLevel& L = g_sys.lv[l];
        L.nx = g.nx >> l; L.ny = g.ny >> l; L.nz = g.nz >> l;
        const size_t n = (size_t)L.nx * L.ny * L.nz;
        L.sdf.assign(n, 0.f); L.uw.assign((size_t)(L.nx + 1) * L.ny * L.nz, 0.f); L.vw.assign((size_t)L.nx * (L.ny + 1) * L.nz, 0.f);
        L.ww.assign((size_t)L.nx * L.ny * (L.nz + 1), 0.f);
        L.center.assign(n, 0.0); L.cr.assign(n, 0.f); L.cu.assign(n, 0.f); L.cf.assign(n, 0.f); L.code.assign(n + 16, 0);
    }
    HostLevel& L0 = g_sys.lv[0];
    for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx; ++i) fluid_sdf0_cell(g, fluid_sdf, L0.sdf.data(), i, j, k);
    for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx + 1; ++i) face_weight0_face<0>(g, collider_sdf, L0.uw.data(), i, j, k);
    for (int k = 0; k < g.nz; ++k) for (int j = 0; j < g.ny + 1; ++j) for (int i = 0; i < g.nx; ++i) face_weight0_face<1>(g, collider_sdf, L0.vw.data(), i, j, k);
    for (int k = 0; k < g.nz + 1; ++k) for (int j = 0; j < g.ny; ++j) for (int i = 0; i < g.nx; ++i) face_weight0_face<2>(g, collider_sdf, L0.ww.data(), i, j, k);
    auto restrict_all = [](const std::vector<float>& f, int fx, int fy, int fz, std::vector<float>& c, int cx, int cy, int cz) {
        for (int k = 0; k < cz; ++k) for (int j = 0; j < cy; ++j) for (int i = 0; i < cx; ++i)
            restrict_f32_cell(f.data(), fx, fy, fz, c.data(), cx, cy, cz, i, j, k);
    };
    for (uint32_t l = 1; l < levels; ++l) {
        HostLevel &F = g_sys.lv[l - 1], &C = g_sys.lv[l];
        restrict_all(F.sdf, F.nx, F.ny, F.nz, C.sdf, C.nx, C.ny, C.nz);
        restrict_all(F.uw, F.nx + 1, F.ny, F.nz, C.uw, C.nx + 1, C.ny, C.nz);
        restrict_all(F.vw, F.nx, F.ny + 1, F.nz, C.vw, C.nx, C.ny + 1, C.nz);
        restrict_all(F.ww, F.nx, F.ny, F.nz + 1, C.ww, C.nx, C.ny, C.nz + 1);
    }
    g_sys.b0.assign((size_t)g.nx * g.ny * g.nz, 0.0);
    double hx = g.hx, hy = g.hy, hz = g.hz;
    for (uint32_t l = 0; l < levels; ++l) {
        HostLevel& L = g_sys.lv[l];
        for (int k = 0; k < L.nz; ++k) for (int j = 0; j < L.ny; ++j) for (int i = 0; i < L.nx; ++i)
            build_system_cell(L.nx, L.ny, L.nz, hx, hy, hz, L.sdf.data(), L.uw.data(), L.vw.data(), L.ww.data(), l == 0 ? u : nullptr,
                              l == 0 ? v : nullptr, l == 0 ? w : nullptr, l == 0 ? bu : nullptr, l == 0 ? bv : nullptr,
                              l == 0 ? bw : nullptr, L.center.data(), L.cr.data(), L.cu.data(), L.cf.data(), L.code.data(),
                              l == 0 ? g_sys.b0.data() : nullptr, i, j, k);
        Mat m;
        m.packed = true;
        m.center = L.center.data(); m.cr = L.cr.data(); m.cu = L.cu.data(); m.cf = L.cf.data(); m.code = L.code.data();
        m.stride_y = (size_t)L.nx; m.stride_z = (size_t)L.nx * L.ny;
        m.sx = (1.0 / hx) * (1.0 / hx); m.sy = (1.0 / hy) * (1.0 / hy); m.sz = (1.0 / hz) * (1.0 / hz);
        {
            volatile double c = 0.0;
            c += m.sx; c += m.sx; c += m.sy; c += m.sy; c += m.sz; c += m.sz;
            m.c6 = c;
        }
        L.m = m;
        hx *= 2.0; hy *= 2.0; hz *= 2.0;
    }
    return (int)levels;
}

// k_mat_to_aos<true> + b0
int hc_pressure_get_rows(uint32_t level, jetcu_matrix_row* A, double* b) {
    if (level >= g_sys.lv.size()) return 1;
    const HostLevel& L = g_sys.lv[level];
    const size_t n = (size_t)L.nx * L.ny * L.nz;
    for (size_t i = 0; i < n; ++i) {
        A[i].center = L.m.center[i]; A[i].right = m_right<true>(L.m, i); A[i].up = m_up<true>(L.m, i); A[i].front = m_front<true>(L.m, i);
    }
    if (b && level == 0) memcpy(b, g_sys.b0.data(), n * 8);
    return 0;
}
int hc_packed_mvm(uint32_t level, const double* v, double* out) {
    const HostLevel& L = g_sys.lv[level];
    for (int k = 0; k < L.nz; ++k) for (int j = 0; j < L.ny; ++j) for (int i = 0; i < L.nx; ++i) mvm_cell<true>(L.m, v, out, i, j, k, L.nx, L.ny, L.nz);
    return 0;
}
int hc_packed_residual(uint32_t level, const double* x, const double* b, double* out) {
    const HostLevel& L = g_sys.lv[level];
    for (int k = 0; k < L.nz; ++k) for (int j = 0; j < L.ny; ++j) for (int i = 0; i < L.nx; ++i) residual_cell<true>(L.m, x, b, out, i, j, k, L.nx, L.ny, L.nz);
    return 0;
}
// use_codes = 1: k_relax_colour_code (what the solver launches); 0: k_relax_colour<true> (float weights, no codes)
int hc_packed_relax(uint32_t level, const double* b, double sor, unsigned iterations, int use_codes, double* x) {
    const HostLevel& L = g_sys.lv[level];
    relax_sweeps<true>(L.m, b, x, sor, iterations, L.nx, L.ny, L.nz, use_codes != 0);
    return 0;
}
// the natural ordering on a packed level (k_relax_lex_tiles<true, true> / k_relax_lex_small<true>)
int hc_packed_relax_lex(uint32_t level, const double* b, double sor, unsigned iterations, int order, double* x) {
    const HostLevel& L = g_sys.lv[level];
    relax_lex_schedule<true>(L.m, b, x, sor, iterations, L.nx, L.ny, L.nz, L.m.code != nullptr, order);
    return 0;
}
// launch_sweep<true> of fdm.cu with JETCU_RELAX_CHUNK = chunk: the red / black z-chunk interleaving, replayed with the
// same loop structure (a "launch" = one colour over a plane range)
int hc_packed_relax_chunked(uint32_t level, const double* b, double sor, unsigned iterations, int chunk, double* x) {
    const HostLevel& L = g_sys.lv[level];
    const int nx = L.nx, ny = L.ny, nz = L.nz;
    auto colour_range = [&](int colour, int k0, int k1) {
        for (int k = k0; k < k1; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = (j + k + colour) & 1; i < nx; i += 2) relax_cell_code(L.m, b, x, sor, i, j, k, nx, ny, nz);
    };
    for (unsigned it = 0; it < iterations; ++it) {
        const int rw0 = 0, rw1 = nz, bw0 = 0, bw1 = nz;
        if (chunk <= 0 || bw1 - bw0 <= chunk) { colour_range(0, rw0, rw1); colour_range(1, bw0, bw1); continue; }
        int red_done = rw0;
        for (int c0 = bw0; c0 < bw1; c0 += chunk) {
            const int c1 = c0 + chunk < bw1 ? c0 + chunk : bw1;
            const int need = c1 + 1 < rw1 ? c1 + 1 : rw1;
            if (need > red_done) { colour_range(0, red_done, need); red_done = need; }
            colour_range(1, c0, c1);
        }
        if (red_done < rw1) colour_range(0, red_done, rw1);
    }
    return 0;
}

// launch_sweeps_skewed<true> of fdm.cu (JETCU_RELAX_SKEW = skew sweeps per group, JETCU_RELAX_CHUNK = chunk planes): the
// skewed wavefront of 2 * skew colour passes over z-chunks, replayed with the same loop structure
int hc_packed_relax_skewed(uint32_t level, const double* b, double sor, unsigned iterations, int chunk, int skew, double* x) {
    const HostLevel& L = g_sys.lv[level];
    const int nx = L.nx, ny = L.ny, nz = L.nz;
    auto colour_range = [&](int colour, int k0, int k1) {
        for (int k = k0; k < k1; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = (j + k + colour) & 1; i < nx; i += 2) relax_cell_code(L.m, b, x, sor, i, j, k, nx, ny, nz);
    };
    unsigned it = 0;
    while (it < iterations) {
        const int s2 = (int)((iterations - it < (unsigned)skew) ? iterations - it : (unsigned)skew);
        const int passes = 2 * s2;
        for (int c0 = 0;; c0 += chunk) {
            bool any = false;
            for (int p = 0; p < passes; ++p) {
                int lo = c0 - p, hi = c0 + chunk - p;
                if (lo < 0) lo = 0;
                if (hi > nz) hi = nz;
                if (c0 - p >= nz) continue;
                any = true;
                if (lo < hi) colour_range(p & 1, lo, hi);
            }
            if (!any) break;
        }
        it += (unsigned)s2;
    }
    return 0;
}

// launch_relax<true>'s slab branch with the wavefront on: one block of k sweeps on the slab [z0, z1) whose 2k halo planes
// are valid in x (here: the whole array is), pass p on the slab +- (2k - 1 - p) planes, skewed over z-chunks.  Only the
// planes [z0, z1) of the result are meaningful.
int hc_packed_relax_skewed_slab(uint32_t level, const double* b, double sor, int k, int chunk, int z0, int z1, double* x) {
    const HostLevel& L = g_sys.lv[level];
    const int nx = L.nx, ny = L.ny, nz = L.nz;
    auto colour_range = [&](int colour, int k0, int k1) {
        for (int kk = k0; kk < k1; ++kk)
            for (int j = 0; j < ny; ++j)
                for (int i = (j + kk + colour) & 1; i < nx; i += 2) relax_cell_code(L.m, b, x, sor, i, j, kk, nx, ny, nz);
    };
    const int passes = 2 * k;
    std::vector<int> w0(passes), w1(passes);
    for (int p = 0; p < passes; ++p) {
        const int e = 2 * k - 1 - p;
        w0[p] = z0 - e < 0 ? 0 : z0 - e; w1[p] = z1 + e > nz ? nz : z1 + e;
    }
    for (int c0 = w0[0];; c0 += chunk) {
        bool any = false;
        for (int p = 0; p < passes; ++p) {
            if (c0 - p >= w1[p]) continue;
            any = true;
            const int lo = c0 - p > w0[p] ? c0 - p : w0[p], hi = c0 + chunk - p < w1[p] ? c0 + chunk - p : w1[p];
            if (lo < hi) colour_range(p & 1, lo, hi);
        }
        if (!any) break;
    }
    return 0;
}

// k_apply_gradient on level 0 of the last build
int hc_apply_gradient(const double* x, const double* u_in, const double* v_in, const double* w_in, double* u_out, double* v_out, double* w_out) {
    const HostLevel& L = g_sys.lv[0];
    const jetcu_grid& g = g_sys.grid;
    for (int k = 0; k < L.nz; ++k) for (int j = 0; j < L.ny; ++j) for (int i = 0; i < L.nx; ++i)
        apply_gradient_cell(L.nx, L.ny, L.nz, g.hx, g.hy, g.hz, L.sdf.data(), L.uw.data(), L.vw.data(), L.ww.data(), x, u_in, v_in, w_in,
                            u_out, v_out, w_out, i, j, k);
    return 0;
}

// diffusion.cu's k_diff_markers + k_diff_system behind plain loops: markers, matrix rows and right-hand side of one array of
// GridBackwardEulerDiffusionSolver3::solve (kind 0 = cell-centred, 1/2/3 = u/v/w), as diffusion_solve_array launches them
int hc_diffusion_system(const jetcu_grid* grid, int kind, const double* f, double coefficient, double dt, int boundary_type,
                        const double* csdf, const double* fsdf, char* markers, jetcu_matrix_row* A, double* b) {
    GridD g = make_grid(grid);
    const DiffArray a = diff_array(g, kind);
    const double tc = dt * coefficient;
    const double cx = tc / (g.hx * g.hx), cy = tc / (g.hy * g.hy), cz = tc / (g.hz * g.hz);
    for (int k = 0; k < a.sz; ++k) for (int j = 0; j < a.sy; ++j) for (int i = 0; i < a.sx; ++i)
        markers[((size_t)k * a.sy + j) * a.sx + i] = diff_marker(g, a, csdf, fsdf, i, j, k);
    for (int k = 0; k < a.sz; ++k) for (int j = 0; j < a.sy; ++j) for (int i = 0; i < a.sx; ++i) {
        const size_t c = ((size_t)k * a.sy + j) * a.sx + i;
        diff_matrix_row(markers, a.sx, a.sy, a.sz, i, j, k, cx, cy, cz, boundary_type == 0, &A[c].center, &A[c].right, &A[c].up, &A[c].front);
        b[c] = diff_rhs(markers, f, a.sx, a.sy, a.sz, i, j, k, cx, cy, cz, boundary_type == 0);
    }
    return 0;
}

}  // extern "C"
