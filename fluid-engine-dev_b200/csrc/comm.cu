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

#include "common.cuh"

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
};

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
    if (c->d_small) cudaFree(c->d_small);
    if (c->h_small) cudaFreeHost(c->h_small);
    delete c;
    ctx->comm = nullptr;
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
    return JETCU_OK;
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
