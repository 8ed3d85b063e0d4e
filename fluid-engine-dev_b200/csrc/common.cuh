// common.cuh — context, error plumbing and small device helpers shared by all jetcu translation units.
// Target: sm_100a only (B200). Built with -fmad=false: the reference is compiled without FMA contraction and
// every stencil here is HBM-bound, so exact IEEE replay of Jet's expressions costs nothing.
#pragma once
#ifndef JETCU_PDL_DEFAULT
#define JETCU_PDL_DEFAULT 1
#endif

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/jetcu.h"

struct jetcu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    std::string error;
    int sm_count = 148;
    // last pressure system (kept for jetcu_pressure_get_system and reused between solves)
    struct PressureSystem* psys = nullptr;
    // reusable PCG workspace
    struct MgWorkspace* mgws = nullptr;
    // pinned staging for scalar read-backs
    double* h_scalars = nullptr;
    double* d_scalars = nullptr;
    double* d_partials = nullptr;  // dot-product partial sums (JC_PARTIALS_CAP doubles)
    // z-slab communicator (comm.cu); null = single GPU
    struct JcComm* comm = nullptr;
    int live_graphs = 0;     // instantiated V-cycle graphs (they pin the NCCL communicator they captured)
    bool graphs_ok = true;   // cleared (for good) the first time a V-cycle capture fails; JETCU_NO_GRAPHS=1 also clears it
    // optional per-kernel-family timing (jetcu_profile_*): CUDA event pairs on `stream`, resolved lazily
    bool profile = false;
    struct ProfPending { cudaEvent_t e0, e1; int kind; uint64_t launches, units; };
    std::vector<ProfPending> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[JETCU_PROF_KINDS] = {};
    uint64_t prof_launches[JETCU_PROF_KINDS] = {};
    uint64_t prof_units[JETCU_PROF_KINDS] = {};
};

#define JC_PARTIALS_CAP 4096

// ---- multi-GPU plumbing (comm.cu): NCCL over NVLink, one rank per GPU; all calls enqueue on ctx->stream ------------
int jc_comm_world(const jetcu_ctx* ctx);
int jc_comm_rank(const jetcu_ctx* ctx);
// Arrays are full-size and globally indexed on every rank; this rank owns planes [k0, k1) of nz.  Sends its `depth`
// lowest planes down / highest planes up and receives planes [k0-depth, k0) / [k1, k1+depth) from the neighbouring
// ranks (no-op at the domain ends; depth must not exceed the slab thickness).
int comm_halo_planes(jetcu_ctx* ctx, void* base, size_t plane_bytes, int k0, int k1, int nz, int depth = 1);
// Peer-memory variant of the same exchange (JETCU_P2P_HALO=1, comm.cu): its bounded waits raise a sticky flag instead of
// hanging; comm_p2p_check reads it (one host sync) and turns it into an error.
int comm_p2p_check(jetcu_ctx* ctx);
// In-place all-gather: rank r contributes bytes [r*bytes_per_rank, (r+1)*bytes_per_rank) of `base`.
int comm_allgather_inplace(jetcu_ctx* ctx, void* base, size_t bytes_per_rank);
int comm_allreduce_max_f64(jetcu_ctx* ctx, double* d_values, size_t n);
int comm_bcast(jetcu_ctx* ctx, void* d_ptr, size_t bytes, int root);
// all-gather of small per-rank records through the device: out[r*count .. ] = rank r's `count` u64 values (host result)
int comm_allgather_u64_host(jetcu_ctx* ctx, const uint64_t* mine, size_t count, uint64_t* all);
// one message per peer: send_ptr[q]/send_bytes[q] to rank q, recv_ptr[q]/recv_bytes[q] from rank q (zero = skip)
int comm_exchange(jetcu_ctx* ctx, void* const* send_ptr, const size_t* send_bytes, void* const* recv_ptr,
                  const size_t* recv_bytes);
void comm_free(jetcu_ctx* ctx);
// message / byte counters (a replayed CUDA graph adds what its capture counted)
void comm_counters(const jetcu_ctx* ctx, uint64_t* messages, uint64_t* bytes);
void comm_counters_add(jetcu_ctx* ctx, uint64_t messages, uint64_t bytes);

// Slab routing rule shared by the device kernel (k_slab_route) and the host-side test hook (jetcu_slab_route_mask):
// bit q is set for every rank q != me whose extended slab [q*T - ghost, (q+1)*T + ghost) contains cell plane ck.
// With T >= 2*ghost only the owner and its two neighbours can qualify.
__host__ __device__ inline unsigned long long jc_slab_dest_mask(int ck, int T, int world, int me, int ghost) {
    unsigned long long m = 0ull;
    const int owner = ck / T;
    for (int dq = -1; dq <= 1; ++dq) {
        const int q = owner + dq;
        if (q < 0 || q >= world || q == me) continue;
        if (ck < q * T - ghost || ck >= (q + 1) * T + ghost) continue;
        m |= 1ull << q;
    }
    return m;
}

// Brackets a group of launches of one kernel family with events when profiling is on (no-ops otherwise).
int jc_prof_begin(jetcu_ctx* ctx, int kind);
void jc_prof_end(jetcu_ctx* ctx, int slot, uint64_t launches, uint64_t units);
int jc_prof_resolve(jetcu_ctx* ctx);

int jetcu_fail(jetcu_ctx* ctx, int code, const char* fmt, ...);

#define JC_CUDA(ctx, expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return jetcu_fail((ctx), JETCU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                          \
    } while (0)

#define JC_TRY(expr)                 \
    do {                             \
        int _rc = (expr);            \
        if (_rc != JETCU_OK) return _rc; \
    } while (0)

#define JC_ARG(ctx, cond)                                                                           \
    do {                                                                                            \
        if (!(cond)) return jetcu_fail((ctx), JETCU_ERR_INVALID_ARG, "invalid argument: %s", #cond); \
    } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched through JC_LAUNCH_PDL may become resident while the previous
// kernel of the stream is still draining — its CTAs call jc_pdl_prologue() first, which (1) lets the NEXT dependent
// launch start as early as possible and (2) blocks until the previous grid has completed and its writes are visible.
// Only the launch gap and the tail of the previous grid are overlapped, never the data: results are unchanged.  The
// kernels that call the prologue behave exactly as before when launched without the attribute (the wait returns at once).
// JETCU_PDL=0 switches the attribute off (plain stream order).
__device__ __forceinline__ void jc_pdl_prologue() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
static inline bool jc_pdl_on() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("JETCU_PDL"); v = e ? (atoi(e) != 0) : JETCU_PDL_DEFAULT; }
    return v != 0;
}
template <typename... KA, typename... CA>
static inline cudaError_t jc_launch_pdl(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, CA&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}
// `pdl` = this launch directly follows another kernel launch of the same stream (not a copy, memset or collective)
#define JC_LAUNCH_PDL(ctx, pdl, kernel, grid, block, smem, ...)                                     \
    do {                                                                                            \
        if ((pdl) && jc_pdl_on()) {                                                                 \
            cudaError_t _l = jc_launch_pdl(kernel, (grid), (block), (smem), (ctx)->stream, __VA_ARGS__); \
            if (_l != cudaSuccess)                                                                  \
                return jetcu_fail((ctx), JETCU_ERR_CUDA, "dependent launch of %s failed: %s (%s:%d)", #kernel, \
                                  cudaGetErrorString(_l), __FILE__, __LINE__);                      \
        } else {                                                                                    \
            kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                        \
        }                                                                                           \
        (ctx)->launches++;                                                                          \
        cudaError_t _e = cudaPeekAtLastError();                                                     \
        if (_e != cudaSuccess)                                                                      \
            return jetcu_fail((ctx), JETCU_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,    \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                          \
    } while (0)

// Every kernel launch goes through this so that jetcu_launch_count() is an honest count.
#define JC_LAUNCH(ctx, kernel, grid, block, smem, ...)                                              \
    do {                                                                                            \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                            \
        (ctx)->launches++;                                                                          \
        cudaError_t _e = cudaPeekAtLastError();                                                     \
        if (_e != cudaSuccess)                                                                      \
            return jetcu_fail((ctx), JETCU_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,    \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                          \
    } while (0)

// Stream-ordered device allocation (cudaMallocAsync pool of the ctx's device).
template <typename T>
static inline int jc_alloc(jetcu_ctx* ctx, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMallocAsync((void**)p, count * sizeof(T), ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return jetcu_fail(ctx, e == cudaErrorMemoryAllocation ? JETCU_ERR_OOM : JETCU_ERR_CUDA,
                          "cudaMallocAsync(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    }
    return JETCU_OK;
}

template <typename T>
static inline void jc_free(jetcu_ctx* ctx, T*& p) {
    if (p) cudaFreeAsync((void*)p, ctx->stream);
    p = nullptr;
}

template <typename T>
static inline int jc_h2d(jetcu_ctx* ctx, T* dst, const T* src, size_t count) {
    JC_CUDA(ctx, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return JETCU_OK;
}

template <typename T>
static inline int jc_d2h(jetcu_ctx* ctx, T* dst, const T* src, size_t count) {
    JC_CUDA(ctx, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return JETCU_OK;
}

template <typename T>
static inline int jc_d2d(jetcu_ctx* ctx, T* dst, const T* src, size_t count) {
    JC_CUDA(ctx, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    return JETCU_OK;
}

template <typename T>
static inline int jc_zero(jetcu_ctx* ctx, T* dst, size_t count) {
    JC_CUDA(ctx, cudaMemsetAsync(dst, 0, count * sizeof(T), ctx->stream));
    return JETCU_OK;
}

static inline int jc_sync(jetcu_ctx* ctx) {
    JC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JETCU_OK;
}

static inline unsigned jc_blocks(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// Device layouts of Jet's FdmMatrix3 (AoS 32-byte rows on the host).
//  * MatSoA:   four contiguous nx*ny*nz double planes (any 7-point matrix; seam 1, host-supplied systems).
//  * packed:   what GridFractionalSinglePhasePressureSolver3::buildSingleSystem can produce — center as double plus
//              the three +face weights as the reference's own floats; an off-diagonal is either 0 or exactly
//              0.0 - double(weight) * invH^2 (…pressure_solver3.cpp:146-150), so it is rebuilt bit-exactly on the fly
//              from 4 bytes instead of being read as 8: 20 B/row instead of 32 B/row.
struct MatSoA {
    double* center = nullptr;
    double* right = nullptr;
    double* up = nullptr;
    double* front = nullptr;
};

struct Mat {
    const double* center = nullptr;
    const double* right = nullptr;  // generic form
    const double* up = nullptr;
    const double* front = nullptr;
    const float* cr = nullptr;      // packed form: +x/+y/+z face weight where the coupling exists, else 0
    const float* cu = nullptr;
    const float* cf = nullptr;
    const uint8_t* code = nullptr;  // packed form: per cell, 2 bits per +face (right | up << 2 | front << 4):
                                    // 0 no coupling, 1 weight == 1.0f, 2 fractional weight (read cr/cu/cf)
    double sx = 0, sy = 0, sz = 0;  // invH^2 per axis (packed form)
    double c6 = 0;                  // packed form: the diagonal of a cell whose six faces are all fully open liquid-liquid
                                    // couplings, ((((((0+sx)+sx)+sy)+sy)+sz)+sz) as buildSingleSystem accumulates it
    size_t stride_y = 0, stride_z = 0;  // nx, nx*ny (packed form; used on the rare fractional-weight path)
    bool packed = false;
};

static inline Mat mat_of(const MatSoA& a) {
    Mat m;
    m.center = a.center; m.right = a.right; m.up = a.up; m.front = a.front;
    return m;
}

struct Dim3i {
    int nx, ny, nz;
    __host__ __device__ size_t n() const { return (size_t)nx * ny * nz; }
};
