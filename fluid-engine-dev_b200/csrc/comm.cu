// comm.cu — z-slab communicator: NCCL over NVLink 5 / NVSwitch, one rank (process) per GPU.
//
// Jet itself has no decomposition at all: everything here is new.  The library resolves NCCL at run
// time (dlopen of libnccl.so.2 — the copy torch already mapped into the process when there is one), so a single-GPU
// user never needs NCCL and libjetcu.so has no link-time dependency on it.  The 128-byte ncclUniqueId travels out of
// band (torch.distributed / MPI / a file): jetcu_comm_unique_id on rank 0, jetcu_comm_init on every rank.
//
// All operations are enqueued on ctx->stream, so they order with the kernels around them without host syncs.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>
#include <stddef.h>
#include <stdio.h>

#include "common.cuh"
#include "p2p_ops.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("JETCU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = std::string("cannot load NCCL: ") + dlerror(); return; }
#define SYM(field, name)                                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));          \
    if (!api.field) { api.error = std::string("NCCL symbol missing: ") + name; return; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
        SYM(GetErrorString, "ncclGetErrorString") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
        SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllGather, "ncclAllGather") SYM(AllReduce, "ncclAllReduce")
        SYM(Broadcast, "ncclBroadcast")
#undef SYM
    });
    return &api;
}

}  // namespace

struct JcComm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    uint64_t* d_small = nullptr;  // staging for comm_allgather_u64_host (world * 64 values)
    uint64_t* h_small = nullptr;
    uint64_t messages = 0, bytes_sent = 0;
    // peer-memory halo path (JETCU_P2P_HALO=1): a mailbox per rank, mapped into both neighbours through CUDA IPC
    bool p2p = false;
    char* box = nullptr;               // [P2P_HEADER | from-below slot 0, 1 | from-above slot 0, 1]
    char* peer_box[2] = {nullptr, nullptr};  // the mailboxes of rank - 1 and rank + 1 in this process's address space
    size_t slot_bytes = 0;
    uint64_t p2p_exchanges = 0;
};

// ---- peer-memory halo exchange -------------------------------------------------------------------------------------------
// NCCL's send/recv pair costs ~27 us per exchange at these sizes and a V-cycle on 8 slabs has ~54 of them:
// latency, not bandwidth.  Here a rank PUSHES its boundary planes straight into the neighbour's mailbox over NVLink
// (plain stores through a CUDA-IPC mapping), publishes a message number behind a system-scope fence, and PULLS its own
// halo out of its mailbox once the neighbours' numbers have arrived: two small kernels, no host involvement, capturable
// in the V-cycle graph (all protocol state lives in device memory).
//   * Two slots per direction, message m uses slot m & 1.  No credit traffic is needed: rank r can only be pushing message
//     m after its pull of m - 1 completed, i.e. after the neighbour pushed m - 1, which (stream order) the neighbour did
//     after finishing its own pull of m - 2 — the previous user of slot m & 1.
//   * Every rank performs the same sequence of exchanges (SPMD), so the per-rank counters agree without negotiation.
//   * Waits are bounded (~2 s of clock64): on expiry a sticky error flag is raised and every later wait returns at once;
//     the solver reads the flag at its next host synchronisation and fails loudly.  Nothing here can hang the GPU.
#define P2P_THREADS 512
#define P2P_TIMEOUT_CYCLES 4000000000ll

__device__ __forceinline__ void p2p_publish(unsigned long long* flag, unsigned long long v) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(flag) = v;
}
__device__ __forceinline__ unsigned long long p2p_peek(const unsigned long long* flag) {
    const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(flag);
    __threadfence_system();
    return v;
}

// push: my boundary planes -> the neighbours' mailboxes; the last CTA to finish publishes the message number
__global__ void __launch_bounds__(P2P_THREADS) k_p2p_push(P2PExchange e) {
    const unsigned long long seq = e.me->seq + 1;
    const size_t off = p2p_slot_offset(e, seq);
    for (size_t i = (size_t)blockIdx.x * P2P_THREADS + threadIdx.x; i < e.n16; i += (size_t)gridDim.x * P2P_THREADS) {
        if (e.src_dn) e.box_dn[off + i] = e.src_dn[i];
        if (e.src_up) e.box_up[off + i] = e.src_up[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&e.me->pushed, 1u);
        if (prev == gridDim.x - 1) {
            e.me->pushed = 0;
            if (e.flag_dn) p2p_publish(e.flag_dn, seq);
            if (e.flag_up) p2p_publish(e.flag_up, seq);
        }
    }
}

// pull: wait for the neighbours' message numbers, copy my halo planes out of my mailbox; the last CTA advances seq
__global__ void __launch_bounds__(P2P_THREADS) k_p2p_pull(P2PExchange e) {
    const unsigned long long seq = e.me->seq + 1;
    const size_t off = p2p_slot_offset(e, seq);
    if (threadIdx.x == 0 && p2p_peek(&e.me->err) == 0) {
        const long long t0 = clock64();
        for (;;) {
            if (p2p_arrived(e, p2p_peek(&e.me->data[0]), p2p_peek(&e.me->data[1]), seq)) break;
            if (clock64() - t0 > P2P_TIMEOUT_CYCLES) { e.me->err = 1; break; }
        }
    }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * P2P_THREADS + threadIdx.x; i < e.n16; i += (size_t)gridDim.x * P2P_THREADS) {
        if (e.dst_dn) e.dst_dn[i] = __ldcg(e.in_dn + off + i);
        if (e.dst_up) e.dst_up[i] = __ldcg(e.in_up + off + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&e.me->pulled, 1u);
        if (prev == gridDim.x - 1) {  // every CTA has read seq (at its start) and finished: safe to advance
            e.me->pulled = 0;
            e.me->seq = seq;
        }
    }
}

static int comm_halo_p2p(jetcu_ctx* ctx, char* p, size_t plane_bytes, int k0, int k1, int depth, bool down, bool up) {
    JcComm* c = ctx->comm;
    const size_t bytes = (size_t)depth * plane_bytes;
    const P2PExchange e = p2p_exchange(c->box, c->peer_box[0], c->peer_box[1], c->slot_bytes, p, plane_bytes, k0, k1, depth, down, up);
    unsigned grid = (unsigned)((e.n16 + (size_t)P2P_THREADS * 4 - 1) / ((size_t)P2P_THREADS * 4));
    if (grid < 1) grid = 1;
    if (grid > 296) grid = 296;
    k_p2p_push<<<grid, P2P_THREADS, 0, ctx->stream>>>(e);
    k_p2p_pull<<<grid, P2P_THREADS, 0, ctx->stream>>>(e);
    ctx->launches += 2;
    JC_CUDA(ctx, cudaGetLastError());
    c->p2p_exchanges += 1;
    c->messages += (down ? 1 : 0) + (up ? 1 : 0);
    c->bytes_sent += ((down ? 1 : 0) + (up ? 1 : 0)) * bytes;
    return JETCU_OK;
}

#define JC_NCCL(ctx, expr)                                                                                    \
    do {                                                                                                      \
        ncclResult_t _r = (expr);                                                                             \
        if (_r != ncclSuccess)                                                                                \
            return jetcu_fail((ctx), JETCU_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString(_r), \
                              __FILE__, __LINE__);                                                            \
    } while (0)

int jc_comm_world(const jetcu_ctx* ctx) { return (ctx && ctx->comm) ? ctx->comm->world : 1; }
int jc_comm_rank(const jetcu_ctx* ctx) { return (ctx && ctx->comm) ? ctx->comm->rank : 0; }

int comm_halo_planes(jetcu_ctx* ctx, void* base, size_t plane_bytes, int k0, int k1, int nz, int depth) {
    JcComm* c = ctx->comm;
    if (!c || c->world <= 1) return JETCU_OK;
    JC_ARG(ctx, depth >= 1 && depth <= k1 - k0);
    NcclApi* api = nccl_api();
    char* p = static_cast<char*>(base);
    const bool down = k0 > 0 && c->rank > 0, up = k1 < nz && c->rank + 1 < c->world;
    if (!down && !up) return JETCU_OK;
    const size_t bytes = (size_t)depth * plane_bytes;
    // the same decision on every rank: sizes and alignment are properties of the level, not of the rank.  The un-owned end
    // ranks (no neighbour on one side) still run both kernels so that their message counters stay in step.
    if (c->p2p && bytes <= c->slot_bytes && bytes % 16 == 0 && plane_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0)
        return comm_halo_p2p(ctx, p, plane_bytes, k0, k1, depth, down, up);
    JC_NCCL(ctx, api->GroupStart());
    if (down) {
        JC_NCCL(ctx, api->Send(p + (size_t)k0 * plane_bytes, bytes, ncclInt8, c->rank - 1, c->comm, ctx->stream));
        JC_NCCL(ctx, api->Recv(p + (size_t)(k0 - depth) * plane_bytes, bytes, ncclInt8, c->rank - 1, c->comm, ctx->stream));
    }
    if (up) {
        JC_NCCL(ctx, api->Send(p + (size_t)(k1 - depth) * plane_bytes, bytes, ncclInt8, c->rank + 1, c->comm, ctx->stream));
        JC_NCCL(ctx, api->Recv(p + (size_t)k1 * plane_bytes, bytes, ncclInt8, c->rank + 1, c->comm, ctx->stream));
    }
    JC_NCCL(ctx, api->GroupEnd());
    c->messages += (down ? 1 : 0) + (up ? 1 : 0);
    c->bytes_sent += ((down ? 1 : 0) + (up ? 1 : 0)) * bytes;
    return JETCU_OK;
}

int comm_allgather_inplace(jetcu_ctx* ctx, void* base, size_t bytes_per_rank) {
    JcComm* c = ctx->comm;
    if (!c || c->world <= 1 || bytes_per_rank == 0) return JETCU_OK;
    char* p = static_cast<char*>(base);
    JC_NCCL(ctx, nccl_api()->AllGather(p + (size_t)c->rank * bytes_per_rank, p, bytes_per_rank, ncclInt8, c->comm, ctx->stream));
    c->messages += 1; c->bytes_sent += bytes_per_rank * (c->world - 1);
    return JETCU_OK;
}

int comm_allreduce_max_f64(jetcu_ctx* ctx, double* d_values, size_t n) {
    JcComm* c = ctx->comm;
    if (!c || c->world <= 1) return JETCU_OK;
    JC_NCCL(ctx, nccl_api()->AllReduce(d_values, d_values, n, ncclDouble, ncclMax, c->comm, ctx->stream));
    c->messages += 1;
    return JETCU_OK;
}

int comm_bcast(jetcu_ctx* ctx, void* d_ptr, size_t bytes, int root) {
    JcComm* c = ctx->comm;
    if (!c || c->world <= 1 || bytes == 0) return JETCU_OK;
    JC_NCCL(ctx, nccl_api()->Broadcast(d_ptr, d_ptr, bytes, ncclInt8, root, c->comm, ctx->stream));
    c->messages += 1;
    return JETCU_OK;
}

int comm_allgather_u64_host(jetcu_ctx* ctx, const uint64_t* mine, size_t count, uint64_t* all) {
    JcComm* c = ctx->comm;
    if (!c || c->world <= 1) { memcpy(all, mine, count * sizeof(uint64_t)); return JETCU_OK; }
    JC_ARG(ctx, count <= 64);
    memcpy(c->h_small + (size_t)c->rank * count, mine, count * sizeof(uint64_t));
    JC_CUDA(ctx, cudaMemcpyAsync(c->d_small + (size_t)c->rank * count, c->h_small + (size_t)c->rank * count,
                                 count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    JC_NCCL(ctx, nccl_api()->AllGather(c->d_small + (size_t)c->rank * count, c->d_small, count, ncclUint64, c->comm, ctx->stream));
    JC_CUDA(ctx, cudaMemcpyAsync(c->h_small, c->d_small, (size_t)c->world * count * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    JC_TRY(jc_sync(ctx));
    memcpy(all, c->h_small, (size_t)c->world * count * sizeof(uint64_t));
    c->messages += 1;
    return JETCU_OK;
}

int comm_exchange(jetcu_ctx* ctx, void* const* send_ptr, const size_t* send_bytes, void* const* recv_ptr,
                  const size_t* recv_bytes) {
    JcComm* c = ctx->comm;
    if (!c || c->world <= 1) return JETCU_OK;
    NcclApi* api = nccl_api();
    JC_NCCL(ctx, api->GroupStart());
    for (int q = 0; q < c->world; ++q) {
        if (q == c->rank) continue;
        if (send_bytes[q]) {
            JC_NCCL(ctx, api->Send(send_ptr[q], send_bytes[q], ncclInt8, q, c->comm, ctx->stream));
            c->messages += 1; c->bytes_sent += send_bytes[q];
        }
        if (recv_bytes[q]) JC_NCCL(ctx, api->Recv(recv_ptr[q], recv_bytes[q], ncclInt8, q, c->comm, ctx->stream));
    }
    JC_NCCL(ctx, api->GroupEnd());
    return JETCU_OK;
}

void comm_counters(const jetcu_ctx* ctx, uint64_t* messages, uint64_t* bytes) {
    *messages = ctx->comm ? ctx->comm->messages : 0;
    *bytes = ctx->comm ? ctx->comm->bytes_sent : 0;
}

void comm_counters_add(jetcu_ctx* ctx, uint64_t messages, uint64_t bytes) {
    if (ctx->comm) { ctx->comm->messages += messages; ctx->comm->bytes_sent += bytes; }
}

void comm_free(jetcu_ctx* ctx) {
    JcComm* c = ctx->comm;
    if (!c) return;
    cudaStreamSynchronize(ctx->stream);
    // NCCL's teardown (destroy and abort alike) polls until every CUDA graph that captured operations of this
    // communicator has been destroyed.  Contexts free their own graphs first (jetcu_ctx_destroy); if the caller still
    // holds solver objects with live V-cycle graphs, the communicator is deliberately leaked instead of hanging here.
    if (c->comm && ctx->live_graphs <= 0) nccl_api()->CommDestroy(c->comm);
    for (char*& pb : c->peer_box)
        if (pb) { cudaIpcCloseMemHandle(pb); pb = nullptr; }
    if (c->box) cudaFree(c->box);
    if (c->d_small) cudaFree(c->d_small);
    if (c->h_small) cudaFreeHost(c->h_small);
    delete c;
    ctx->comm = nullptr;
}

// Sticky time-out flag of the peer-memory path (see P2PFlags): read where the solver synchronises with the host anyway.
int comm_p2p_check(jetcu_ctx* ctx) {
    JcComm* c = ctx->comm;
    if (!c || !c->p2p) return JETCU_OK;
    unsigned long long err = 0;
    JC_CUDA(ctx, cudaMemcpyAsync(&err, c->box + offsetof(P2PFlags, err), sizeof(err), cudaMemcpyDeviceToHost, ctx->stream));
    JC_TRY(jc_sync(ctx));
    if (err) return jetcu_fail(ctx, JETCU_ERR_NCCL, "peer-memory halo exchange: a neighbour's planes did not arrive within the "
                                                    "time-out (rank %d of %d)", c->rank, c->world);
    return JETCU_OK;
}

// JETCU_P2P_HALO=1: allocate the mailbox, swap CUDA-IPC handles through the NCCL communicator, map both neighbours, prove the
// path with one self-checking exchange.  The path is switched on only if EVERY rank got through (an all-reduce decides), so
// the ranks can never disagree about the protocol; otherwise a line on stderr and NCCL send/recv stays in charge.
static int comm_p2p_setup(jetcu_ctx* ctx) {
    JcComm* c = ctx->comm;
    const char* want = getenv("JETCU_P2P_HALO");
    if (!want || atoi(want) == 0) return JETCU_OK;
    const char* mb = getenv("JETCU_P2P_SLOT_MB");
    c->slot_bytes = (size_t)((mb && atoi(mb) > 0) ? atoi(mb) : 16) << 20;
    double fail = 0.0;
    cudaIpcMemHandle_t* d_handles = nullptr;
    std::vector<cudaIpcMemHandle_t> handles((size_t)c->world);
    if (cudaMalloc((void**)&c->box, P2P_HEADER + 4 * c->slot_bytes) != cudaSuccess) { c->box = nullptr; fail = 1.0; }
    if (cudaMalloc((void**)&d_handles, (size_t)c->world * sizeof(cudaIpcMemHandle_t)) != cudaSuccess) { d_handles = nullptr; fail = 1.0; }
    if (fail == 0.0) {
        cudaMemsetAsync(c->box, 0, P2P_HEADER, ctx->stream);
        cudaMemsetAsync(d_handles, 0, (size_t)c->world * sizeof(cudaIpcMemHandle_t), ctx->stream);
        cudaIpcMemHandle_t mine;
        if (cudaIpcGetMemHandle(&mine, c->box) != cudaSuccess) fail = 1.0;
        else cudaMemcpyAsync(d_handles + c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream);
    }
    // every rank takes part in the collectives below whatever happened above
    if (d_handles) {
        if (comm_allgather_inplace(ctx, d_handles, sizeof(cudaIpcMemHandle_t)) != JETCU_OK) fail = 1.0;
        if (cudaMemcpyAsync(handles.data(), d_handles, handles.size() * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost,
                            ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) fail = 1.0;
    }
    if (fail == 0.0) {
        if (c->rank > 0 && cudaIpcOpenMemHandle((void**)&c->peer_box[0], handles[(size_t)c->rank - 1],
                                                cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { c->peer_box[0] = nullptr; fail = 1.0; }
        if (c->rank + 1 < c->world && cudaIpcOpenMemHandle((void**)&c->peer_box[1], handles[(size_t)c->rank + 1],
                                                           cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { c->peer_box[1] = nullptr; fail = 1.0; }
    }
    cudaGetLastError();  // a failed attempt must not poison later launches
    // agreement no. 1: mailboxes mapped everywhere?
    double* d_flag = reinterpret_cast<double*>(c->d_small);
    cudaMemcpyAsync(d_flag, &fail, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    JC_TRY(comm_allreduce_max_f64(ctx, d_flag, 1));
    cudaMemcpyAsync(&fail, d_flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    JC_TRY(jc_sync(ctx));
    if (fail == 0.0) {
        // self-check: 4 planes of 4 KiB per rank, plane value = 16 * rank + plane; three exchanges (both slots and a reuse)
        const size_t plane = 4096;
        std::vector<unsigned char> h(4 * plane);
        unsigned char* w = nullptr;
        if (cudaMalloc((void**)&w, 4 * plane) != cudaSuccess) fail = 1.0;
        c->p2p = true;
        for (int round = 0; round < 3 && fail == 0.0; ++round) {
            for (int q = 0; q < 4; ++q) memset(h.data() + (size_t)q * plane, 16 * c->rank + q + 64 * round, plane);
            cudaMemcpyAsync(w, h.data(), 4 * plane, cudaMemcpyHostToDevice, ctx->stream);
            if (comm_halo_planes(ctx, w, plane, c->rank > 0 ? 1 : 0, c->rank + 1 < c->world ? 3 : 4, 4, 1) != JETCU_OK) fail = 1.0;
            cudaMemcpyAsync(h.data(), w, 4 * plane, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) fail = 1.0;
            // my ghost-lo plane [0] must hold the lower neighbour's own-hi plane (its plane 2), my ghost-hi [3] the upper one's plane 1
            if (c->rank > 0 && (h[0] != (unsigned char)(16 * (c->rank - 1) + 2 + 64 * round) || h[plane - 1] != h[0])) fail = 1.0;
            if (c->rank + 1 < c->world && (h[3 * plane] != (unsigned char)(16 * (c->rank + 1) + 1 + 64 * round) || h[4 * plane - 1] != h[3 * plane])) fail = 1.0;
        }
        if (comm_p2p_check(ctx) != JETCU_OK) fail = 1.0;
        if (w) cudaFree(w);
        cudaGetLastError();
        // agreement no. 2: did the data arrive everywhere?
        cudaMemcpyAsync(d_flag, &fail, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        c->p2p = false;  // the vote itself goes through NCCL
        JC_TRY(comm_allreduce_max_f64(ctx, d_flag, 1));
        cudaMemcpyAsync(&fail, d_flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        JC_TRY(jc_sync(ctx));
    }
    if (d_handles) cudaFree(d_handles);
    c->p2p = fail == 0.0;
    if (!c->p2p) {
        fprintf(stderr, "jetcu: JETCU_P2P_HALO requested but the peer-memory path is unavailable on rank %d; using NCCL send/recv\n", c->rank);
        for (char*& pb : c->peer_box)
            if (pb) { cudaIpcCloseMemHandle(pb); pb = nullptr; }
        if (c->box) { cudaFree(c->box); c->box = nullptr; }
        cudaGetLastError();
        ctx->error.clear();
    }
    return JETCU_OK;
}

extern "C" {

int jetcu_comm_unique_id(uint8_t id[JETCU_UNIQUE_ID_BYTES]) {
    NcclApi* api = nccl_api();
    if (!api->lib || !api->GetUniqueId) return jetcu_fail(nullptr, JETCU_ERR_NCCL, "%s", api->error.c_str());
    static_assert(sizeof(ncclUniqueId) == JETCU_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    ncclResult_t r = api->GetUniqueId(&u);
    if (r != ncclSuccess) return jetcu_fail(nullptr, JETCU_ERR_NCCL, "ncclGetUniqueId: %s", api->GetErrorString(r));
    memcpy(id, &u, sizeof(u));
    return JETCU_OK;
}

int jetcu_comm_init(jetcu_ctx* ctx, int world, int rank, const uint8_t id[JETCU_UNIQUE_ID_BYTES]) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    JC_ARG(ctx, world >= 1 && rank >= 0 && rank < world && id != nullptr);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");
    comm_free(ctx);
    if (world == 1) return JETCU_OK;
    NcclApi* api = nccl_api();
    if (!api->lib || !api->Broadcast) return jetcu_fail(ctx, JETCU_ERR_NCCL, "%s", api->error.c_str());
    JcComm* c = new JcComm();
    c->world = world; c->rank = rank;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return jetcu_fail(ctx, JETCU_ERR_NCCL, "ncclCommInitRank(world %d, rank %d): %s", world, rank, api->GetErrorString(r));
    }
    if (cudaMalloc((void**)&c->d_small, (size_t)world * 64 * sizeof(uint64_t)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_small, (size_t)world * 64 * sizeof(uint64_t)) != cudaSuccess) {
        api->CommDestroy(c->comm);
        delete c;
        return jetcu_fail(ctx, JETCU_ERR_CUDA, "communicator staging allocation failed");
    }
    ctx->comm = c;
    // NCCL connects peers / rings lazily at the first operation that needs them, and that set-up is not legal inside a
    // stream capture.  The V-cycle is replayed from a CUDA graph whose capture may well be the first traffic on this
    // communicator, so every pattern the graphs contain is exercised once, eagerly, right here: halo send/recv with both
    // neighbours (small and halo-sized, so that all p2p channels come up), all-gather, all-reduce and broadcast.
    {
        const size_t big = (size_t)32 << 20;
        char* w = nullptr;
        if (cudaMalloc((void**)&w, 4 * big) != cudaSuccess) { comm_free(ctx); return jetcu_fail(ctx, JETCU_ERR_OOM, "communicator warm-up buffer"); }
        cudaMemsetAsync(w, 0, 4 * big, ctx->stream);
        int rc = JETCU_OK;
        for (size_t bytes : {(size_t)4096, big}) {
            // planes of `bytes` each: [0] ghost-lo, [1] own-lo, [2] own-hi, [3] ghost-hi  ->  k0 = 1, k1 = 3 of nz = 4
            if (rc == JETCU_OK) rc = comm_halo_planes(ctx, w, bytes, rank > 0 ? 1 : 0, rank + 1 < world ? 3 : 4, 4, 1);
            if (rc == JETCU_OK) rc = comm_allgather_inplace(ctx, w, bytes / (size_t)world);
        }
        if (rc == JETCU_OK) rc = comm_allreduce_max_f64(ctx, reinterpret_cast<double*>(w), 1);
        if (rc == JETCU_OK) rc = comm_bcast(ctx, w, 4096, world - 1);
        if (rc == JETCU_OK) rc = jc_sync(ctx);
        cudaFree(w);
        if (rc != JETCU_OK) { std::string keep = ctx->error; comm_free(ctx); ctx->error = keep; return rc; }
        c->messages = 0; c->bytes_sent = 0;
    }
    {
        const int rc = comm_p2p_setup(ctx);
        if (rc != JETCU_OK) { std::string keep = ctx->error; comm_free(ctx); ctx->error = keep; return rc; }
        c->messages = 0; c->bytes_sent = 0; c->p2p_exchanges = 0;
    }
    return JETCU_OK;
}

int jetcu_comm_p2p_enabled(const jetcu_ctx* ctx) { return (ctx && ctx->comm && ctx->comm->p2p) ? 1 : 0; }

// Diagnostic: `iters` back-to-back halo exchanges of `depth` planes of `plane_bytes` on a scratch slab, timed with CUDA events
// on the solver stream — the per-exchange latency that bounds the distributed V-cycle, for whichever transport
// is active (NCCL send/recv or the peer-memory mailboxes).
int jetcu_comm_halo_benchmark(jetcu_ctx* ctx, uint64_t plane_bytes, int depth, int iters, double* ms_per_exchange) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    JC_ARG(ctx, ctx->comm != nullptr && plane_bytes >= 16 && depth >= 1 && depth <= 64 && iters >= 1 && ms_per_exchange != nullptr);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return jetcu_fail(ctx, JETCU_ERR_CUDA, "cudaSetDevice failed");
    JcComm* c = ctx->comm;
    const int nzl = 4 * depth;                          // [ghost-lo | own (2 * depth planes) | ghost-hi]
    char* w = nullptr;
    JC_CUDA(ctx, cudaMalloc((void**)&w, (size_t)nzl * plane_bytes));
    cudaMemsetAsync(w, 0, (size_t)nzl * plane_bytes, ctx->stream);
    const int k0 = c->rank > 0 ? depth : 0, k1 = c->rank + 1 < c->world ? 3 * depth : nzl;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = JETCU_OK;
    for (int i = 0; i < 5 && rc == JETCU_OK; ++i) rc = comm_halo_planes(ctx, w, (size_t)plane_bytes, k0, k1, nzl, depth);
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < iters && rc == JETCU_OK; ++i) rc = comm_halo_planes(ctx, w, (size_t)plane_bytes, k0, k1, nzl, depth);
    cudaEventRecord(e1, ctx->stream);
    if (rc == JETCU_OK) rc = jc_sync(ctx);
    if (rc == JETCU_OK) rc = comm_p2p_check(ctx);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(w);
    *ms_per_exchange = (double)ms / (double)iters;
    return rc;
}

int jetcu_comm_destroy(jetcu_ctx* ctx) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    comm_free(ctx);
    return JETCU_OK;
}

int jetcu_comm_world(const jetcu_ctx* ctx) { return jc_comm_world(ctx); }
int jetcu_comm_rank(const jetcu_ctx* ctx) { return jc_comm_rank(ctx); }

int jetcu_comm_stats(const jetcu_ctx* ctx, uint64_t* messages, uint64_t* bytes_sent) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    if (messages) *messages = ctx->comm ? ctx->comm->messages : 0;
    if (bytes_sent) *bytes_sent = ctx->comm ? ctx->comm->bytes_sent : 0;
    return JETCU_OK;
}

uint64_t jetcu_slab_route_mask(int64_t cell_plane, uint64_t nz, int world, int me) {
    if (world < 1 || nz % (uint64_t)world != 0 || cell_plane < 0 || (uint64_t)cell_plane >= nz) return 0;
    return jc_slab_dest_mask((int)cell_plane, (int)(nz / (uint64_t)world), world, me, JETCU_SLAB_GHOST_PLANES);
}

int jetcu_slab_range(uint64_t nz, int world, int rank, uint64_t* z0, uint64_t* z1) {
    if (world < 1 || rank < 0 || rank >= world || nz % (uint64_t)world != 0) return JETCU_ERR_INVALID_ARG;
    const uint64_t t = nz / (uint64_t)world;
    if (z0) *z0 = (uint64_t)rank * t;
    if (z1) *z1 = (uint64_t)(rank + 1) * t;
    return JETCU_OK;
}

}  // extern "C"
