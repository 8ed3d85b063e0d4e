// particles.cu — prefix scan, stable LSD radix sort and the hash-grid binning entry point
// (PointParallelHashGridSearcher3::build, src/jet/point_parallel_hash_grid_searcher3.cpp:52-146).
// All integer work: results are bit-exact against the reference by construction (stable sort == ascending
// original index inside a bucket == PointHashGridSearcher3::buckets() order).
#include "particles.cuh"

#include <math.h>

// ------------------------------------------------------------------------------------------------ scan
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t btotal;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < SCAN_THREADS / 32) ? wsum[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        if (lane < SCAN_THREADS / 32) wsum[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1) btotal = wi;
    }
    __syncthreads();
    uint32_t r = incl - v + wsum[warp];
    *total = btotal;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in,
                                                              uint32_t* __restrict__ sums, size_t n) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t acc = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q)
        if (base + q < n) acc += in[base + q];
    uint32_t total;
    block_excl_scan(acc, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* in, uint32_t* out,
                                                             const uint32_t* __restrict__ sums, size_t n) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t acc = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        v[q] = (base + q < n) ? in[base + q] : 0;
        acc += v[q];
    }
    uint32_t total;
    uint32_t prefix = block_excl_scan(acc, &total) + (sums ? sums[blockIdx.x] : 0);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        if (base + q < n) out[base + q] = prefix;
        prefix += v[q];
    }
}

int scan_exclusive_u32(jetcu_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n) {
    if (n == 0) return JETCU_OK;
    unsigned blocks = (unsigned)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (blocks == 1) {
        JC_LAUNCH(ctx, k_scan_apply, 1, SCAN_THREADS, 0, in, out, (const uint32_t*)nullptr, n);
        return JETCU_OK;
    }
    uint32_t* sums = nullptr;
    JC_TRY(jc_alloc(ctx, &sums, blocks));
    JC_LAUNCH(ctx, k_scan_reduce, blocks, SCAN_THREADS, 0, in, sums, n);
    int rc = scan_exclusive_u32(ctx, sums, sums, blocks);
    if (rc == JETCU_OK) {
        k_scan_apply<<<blocks, SCAN_THREADS, 0, ctx->stream>>>(in, out, sums, n);
        ctx->launches++;
        if (cudaPeekAtLastError() != cudaSuccess) rc = jetcu_fail(ctx, JETCU_ERR_CUDA, "k_scan_apply launch failed");
    }
    jc_free(ctx, sums);
    return rc;
}

// ------------------------------------------------------------------------------------------------ radix sort
#define RS_THREADS 256
#define RS_CHUNKS 16
#define RS_TILE (RS_THREADS * RS_CHUNKS)
#define RS_WARPS (RS_THREADS / 32)

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint32_t* __restrict__ keys, uint32_t* __restrict__ hist,
                                                           size_t n, int shift, unsigned nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
    for (int c = 0; c < RS_CHUNKS; ++c) {
        size_t i = base + (size_t)c * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// Stable scatter: items of a tile are ranked chunk by chunk (chunk = one item per thread, thread order = item order),
// inside a chunk by warp (match_any) then across warps through per-warp digit counts.
__global__ void __launch_bounds__(RS_THREADS)
    k_radix_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t* __restrict__ keys_out,
                    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ offsets, size_t n, int shift,
                    unsigned nblocks) {
    __shared__ uint32_t base[256];
    __shared__ uint32_t wcount[RS_WARPS][256];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = offsets[(size_t)threadIdx.x * nblocks + blockIdx.x];
    size_t tile = (size_t)blockIdx.x * RS_TILE;
    for (int c = 0; c < RS_CHUNKS; ++c) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) wcount[w][threadIdx.x] = 0;
        __syncthreads();
        size_t i = tile + (size_t)c * RS_THREADS + threadIdx.x;
        bool valid = i < n;
        uint32_t key = valid ? keys[i] : 0xffffffffu;
        uint32_t digit = valid ? ((key >> shift) & 255u) : 256u;  // 256 = "no item": matches only other invalid lanes
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) wcount[warp][digit] = __popc(peers);
        __syncthreads();
        uint32_t pos = 0;
        if (valid) {
            pos = base[digit] + rank_in_warp;
            for (unsigned w = 0; w < warp; ++w) pos += wcount[w][digit];
        }
        __syncthreads();
        // thread d advances digit d's running base by this chunk's count
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) add += wcount[w][threadIdx.x];
        base[threadIdx.x] += add;
        if (valid) {
            keys_out[pos] = key;
            vals_out[pos] = vals[i];
        }
        __syncthreads();
    }
}

int radix_sort_pairs_u32(jetcu_ctx* ctx, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                         uint32_t* vals_out, size_t n, int key_bits) {
    if (n == 0) return JETCU_OK;
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    unsigned nblocks = (unsigned)((n + RS_TILE - 1) / RS_TILE);
    uint32_t* hist = nullptr;
    JC_TRY(jc_alloc(ctx, &hist, (size_t)256 * nblocks));
    uint32_t *ki = keys_in, *vi = vals_in, *ko = keys_out, *vo = vals_out;
    int rc = JETCU_OK;
    for (int p = 0; p < passes && rc == JETCU_OK; ++p) {
        k_radix_hist<<<nblocks, RS_THREADS, 0, ctx->stream>>>(ki, hist, n, 8 * p, nblocks);
        ctx->launches++;
        rc = scan_exclusive_u32(ctx, hist, hist, (size_t)256 * nblocks);
        if (rc != JETCU_OK) break;
        k_radix_scatter<<<nblocks, RS_THREADS, 0, ctx->stream>>>(ki, vi, ko, vo, hist, n, 8 * p, nblocks);
        ctx->launches++;
        if (cudaPeekAtLastError() != cudaSuccess) rc = jetcu_fail(ctx, JETCU_ERR_CUDA, "radix pass launch failed");
        uint32_t* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    // after the loop the latest result is in (ki, vi); make sure it ends in the *_out buffers
    if (rc == JETCU_OK && ki != keys_out) {
        rc = jc_d2d(ctx, keys_out, ki, n);
        if (rc == JETCU_OK) rc = jc_d2d(ctx, vals_out, vi, n);
    }
    jc_free(ctx, hist);
    return rc;
}

// ------------------------------------------------------------------------------------------------ hash-grid binning
// getBucketIndex / getHashKeyFromBucketIndex, src/jet/point_parallel_hash_grid_searcher3.cpp:232-269:
// floor(x / spacing) (FP64 division) -> ssize_t; C++ `%` then +res if negative.
__device__ __forceinline__ long long wrap_idx(long long b, long long res) {
    long long w = b % res;
    if (w < 0) w += res;
    return w;
}

__global__ void k_hash_keys(const double* __restrict__ pos, size_t n, double spacing, long long rx, long long ry,
                            long long rz, uint32_t* __restrict__ keys32, uint32_t* __restrict__ idx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long bx = (long long)floor(pos[3 * i] / spacing);
    long long by = (long long)floor(pos[3 * i + 1] / spacing);
    long long bz = (long long)floor(pos[3 * i + 2] / spacing);
    long long key = (wrap_idx(bz, rz) * ry + wrap_idx(by, ry)) * rx + wrap_idx(bx, rx);
    keys32[i] = (uint32_t)key;
    idx[i] = (uint32_t)i;
}

// start/end tables from the sorted keys (:114-125); empty buckets keep kMaxSize.
__global__ void k_bucket_tables(const uint32_t* __restrict__ keys_sorted, size_t n, unsigned long long* __restrict__ start,
                                unsigned long long* __restrict__ end) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = keys_sorted[i];
    if (i == 0) start[k] = 0;
    if (i == n - 1) end[k] = n;
    if (i > 0) {
        uint32_t kp = keys_sorted[i - 1];
        if (k > kp) { start[k] = i; end[kp] = i; }
    }
}

__global__ void k_fill_u64(unsigned long long* p, unsigned long long v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_widen_u32(const uint32_t* __restrict__ in, unsigned long long* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

extern "C" int jetcu_bin_points(jetcu_ctx* ctx, uint64_t n, const double* positions, uint64_t rx, uint64_t ry,
                                uint64_t rz, double spacing, uint64_t* keys_unsorted, uint64_t* sorted_indices,
                                uint64_t* keys_sorted, uint64_t* start_table, uint64_t* end_table) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");
    JC_ARG(ctx, rx > 0 && ry > 0 && rz > 0 && spacing > 0);
    JC_ARG(ctx, n == 0 || positions != nullptr);
    uint64_t nb = rx * ry * rz;
    JC_ARG(ctx, nb < (1ull << 32) && n < (1ull << 32));
    int bits = 1;
    while ((1ull << bits) < nb) ++bits;

    double* dpos = nullptr;
    uint32_t *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr;
    unsigned long long *st = nullptr, *en = nullptr, *wide = nullptr;
    int rc = JETCU_OK;
#define STEP(x) if (rc == JETCU_OK) rc = (x)
    STEP(jc_alloc(ctx, &st, nb)); STEP(jc_alloc(ctx, &en, nb));
    if (rc == JETCU_OK) {
        k_fill_u64<<<jc_blocks(nb, 256), 256, 0, ctx->stream>>>(st, ~0ull, nb);
        k_fill_u64<<<jc_blocks(nb, 256), 256, 0, ctx->stream>>>(en, ~0ull, nb);
        ctx->launches += 2;
    }
    if (n > 0) {
        STEP(jc_alloc(ctx, &dpos, 3 * n)); STEP(jc_h2d(ctx, dpos, positions, 3 * n));
        STEP(jc_alloc(ctx, &k0, n)); STEP(jc_alloc(ctx, &k1, n)); STEP(jc_alloc(ctx, &v0, n)); STEP(jc_alloc(ctx, &v1, n));
        STEP(jc_alloc(ctx, &wide, n));
        if (rc == JETCU_OK) {
            k_hash_keys<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(dpos, n, spacing, (long long)rx, (long long)ry,
                                                                    (long long)rz, k0, v0);
            ctx->launches++;
        }
        if (keys_unsorted && rc == JETCU_OK) {
            k_widen_u32<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(k0, wide, n);
            ctx->launches++;
            STEP(jc_d2h(ctx, (unsigned long long*)keys_unsorted, wide, n));
        }
        STEP(radix_sort_pairs_u32(ctx, k0, v0, k1, v1, n, bits));
        if (rc == JETCU_OK) {
            k_bucket_tables<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(k1, n, st, en);
            ctx->launches++;
        }
        if (sorted_indices && rc == JETCU_OK) {
            k_widen_u32<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(v1, wide, n);
            ctx->launches++;
            STEP(jc_d2h(ctx, (unsigned long long*)sorted_indices, wide, n));
        }
        if (keys_sorted && rc == JETCU_OK) {
            k_widen_u32<<<jc_blocks(n, 256), 256, 0, ctx->stream>>>(k1, wide, n);
            ctx->launches++;
            STEP(jc_d2h(ctx, (unsigned long long*)keys_sorted, wide, n));
        }
    }
    if (start_table) STEP(jc_d2h(ctx, (unsigned long long*)start_table, st, nb));
    if (end_table) STEP(jc_d2h(ctx, (unsigned long long*)end_table, en, nb));
    STEP(jc_sync(ctx));
#undef STEP
    jc_free(ctx, dpos); jc_free(ctx, k0); jc_free(ctx, k1); jc_free(ctx, v0); jc_free(ctx, v1);
    jc_free(ctx, st); jc_free(ctx, en); jc_free(ctx, wide);
    return rc;
}
