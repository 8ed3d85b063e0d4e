// abi.cu — the extern "C" boundary (include/jetcu.h): context management, linear-system seam (1) and
// pressure-projection seam (2).  Host buffers in, host buffers out; all compute on the ctx's stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "fdm.cuh"
#include "pressure.cuh"

static thread_local std::string g_last_create_error;

int jetcu_fail(jetcu_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf; else g_last_create_error = buf;
    return code;
}

int jc_prof_begin(jetcu_ctx* ctx, int kind) {
    if (!ctx->profile) return -1;
    if (ctx->prof_pending.size() >= 4096 && jc_prof_resolve(ctx) != JETCU_OK) return -1;
    jetcu_ctx::ProfPending p{};
    for (cudaEvent_t* e : {&p.e0, &p.e1}) {
        if (!ctx->prof_free.empty()) { *e = ctx->prof_free.back(); ctx->prof_free.pop_back(); }
        else if (cudaEventCreate(e) != cudaSuccess) return -1;
    }
    p.kind = kind;
    cudaEventRecord(p.e0, ctx->stream);
    ctx->prof_pending.push_back(p);
    return (int)ctx->prof_pending.size() - 1;
}

void jc_prof_end(jetcu_ctx* ctx, int slot, uint64_t launches, uint64_t units) {
    if (slot < 0) return;
    jetcu_ctx::ProfPending& p = ctx->prof_pending[slot];
    p.launches = launches; p.units = units;
    cudaEventRecord(p.e1, ctx->stream);
}

int jc_prof_resolve(jetcu_ctx* ctx) {
    if (ctx->prof_pending.empty()) return JETCU_OK;
    JC_TRY(jc_sync(ctx));
    for (auto& p : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            ctx->prof_ms[p.kind] += ms; ctx->prof_launches[p.kind] += p.launches; ctx->prof_units[p.kind] += p.units;
        }
        ctx->prof_free.push_back(p.e0); ctx->prof_free.push_back(p.e1);
    }
    ctx->prof_pending.clear();
    return JETCU_OK;
}

extern "C" {

const char* jetcu_version(void) { return "jetcu 0.1 (sm_100a, fmad=false)"; }

int jetcu_ctx_create(int device, jetcu_ctx** out) {
    if (!out) return jetcu_fail(nullptr, JETCU_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return jetcu_fail(nullptr, JETCU_ERR_CUDA, "no CUDA device available (%s); jetcu has no CPU fallback",
                          e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return jetcu_fail(nullptr, JETCU_ERR_INVALID_ARG, "device %d out of range", device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return jetcu_fail(nullptr, JETCU_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return jetcu_fail(nullptr, JETCU_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                          prop.major, prop.minor);
    jetcu_ctx* ctx = new jetcu_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->h_scalars, 64 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_scalars, 64 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_partials, 4096 * sizeof(double));
    if (e == cudaSuccess) {
        // keep freed blocks in the pool: the solver reallocates the same work vectors every call
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thresh = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
        }
    }
    if (e != cudaSuccess) {
        int rc = jetcu_fail(nullptr, JETCU_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    ctx->psys = new PressureSystem();
    if (getenv("JETCU_NO_GRAPHS")) ctx->graphs_ok = false;
    *out = ctx;
    return JETCU_OK;
}

int jetcu_ctx_destroy(jetcu_ctx* ctx) {
    if (!ctx) return JETCU_OK;
    cudaSetDevice(ctx->device);
    // graphs first: a captured V-cycle pins the NCCL communicator, and ncclCommDestroy waits for such references
    if (ctx->psys) { pressure_free(ctx, ctx->psys); delete ctx->psys; ctx->psys = nullptr; }
    comm_free(ctx);
    cudaStreamSynchronize(ctx->stream);
    jc_prof_resolve(ctx);
    for (cudaEvent_t e : ctx->prof_free) cudaEventDestroy(e);
    cudaFreeHost(ctx->h_scalars);
    cudaFree(ctx->d_scalars);
    cudaFree(ctx->d_partials);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return JETCU_OK;
}

const char* jetcu_last_error(const jetcu_ctx* ctx) { return ctx ? ctx->error.c_str() : g_last_create_error.c_str(); }
uint64_t jetcu_launch_count(const jetcu_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* jetcu_stream(const jetcu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int jetcu_synchronize(jetcu_ctx* ctx) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    return jc_sync(ctx);
}

int jetcu_profile_enable(jetcu_ctx* ctx, int on) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    JC_TRY(jc_prof_resolve(ctx));
    ctx->profile = on != 0;
    if (on) {
        for (int k = 0; k < JETCU_PROF_KINDS; ++k) { ctx->prof_ms[k] = 0; ctx->prof_launches[k] = 0; ctx->prof_units[k] = 0; }
    }
    return JETCU_OK;
}

int jetcu_profile_get(jetcu_ctx* ctx, int kind, double* total_ms, uint64_t* launches, uint64_t* units) {
    if (!ctx) return JETCU_ERR_INVALID_ARG;
    JC_ARG(ctx, kind >= 0 && kind < JETCU_PROF_KINDS);
    JC_TRY(jc_prof_resolve(ctx));
    if (total_ms) *total_ms = ctx->prof_ms[kind];
    if (launches) *launches = ctx->prof_launches[kind];
    if (units) *units = ctx->prof_units[kind];
    return JETCU_OK;
}

uint32_t jetcu_mg_num_levels(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t max_levels) {
    // FdmMgUtils3::resizeArrayWithFinest, include/jet/detail/fdm_mg_linear_system3-inl.h:34-50
    uint32_t i = 1;
    for (; i < max_levels; ++i) {
        if (nx % 2 == 0 && ny % 2 == 0 && nz % 2 == 0) { nx >>= 1; ny >>= 1; nz >>= 1; }
        else break;
    }
    return i;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- helpers
namespace {

struct DevVec {
    jetcu_ctx* ctx;
    double* p = nullptr;
    explicit DevVec(jetcu_ctx* c) : ctx(c) {}
    ~DevVec() { jc_free(ctx, p); }
    int upload(const double* h, size_t n) {
        JC_TRY(jc_alloc(ctx, &p, n));
        if (h) return jc_h2d(ctx, p, h, n);
        return JETCU_OK;
    }
};

struct DevMat {
    jetcu_ctx* ctx;
    MatSoA A;
    explicit DevMat(jetcu_ctx* c) : ctx(c) {}
    ~DevMat() { fdm_free_mat(ctx, &A); }
    int upload(const jetcu_matrix_row* h, size_t n) {
        jetcu_matrix_row* aos = nullptr;
        JC_TRY(jc_alloc(ctx, &aos, n));
        int rc = jc_h2d(ctx, aos, h, n);
        if (rc == JETCU_OK) rc = fdm_alloc_mat(ctx, &A, n);
        if (rc == JETCU_OK) rc = fdm_aos_to_soa(ctx, aos, A, n);
        jc_free(ctx, aos);
        return rc;
    }
};

// slab mode: every rank ends up with the whole level-0 vector
int gather_level0(jetcu_ctx* ctx, MgLevels* mg, double* v) {
    const SlabPlan& pl = mg->plan;
    if (pl.dist_levels == 0) return JETCU_OK;
    const Dim3i d = mg->dims[0];
    return comm_allgather_inplace(ctx, v, (size_t)(pl.z1[0] - pl.z0[0]) * d.nx * d.ny * sizeof(double));
}

int check_dims(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz) {
    JC_ARG(ctx, nx > 0 && ny > 0 && nz > 0);
    JC_ARG(ctx, nx < (1u << 15) && ny < (1u << 15) && nz < (1u << 15));
    return JETCU_OK;
}

int setup_levels(jetcu_ctx* ctx, MgLevels* mg, std::vector<DevMat*>& mats, uint64_t nx, uint64_t ny, uint64_t nz,
                 uint32_t levels, const jetcu_matrix_row* const* A) {
    JC_ARG(ctx, levels >= 1 && levels <= JC_MAX_LEVELS);
    mg->levels = (int)levels;
    for (uint32_t l = 0; l < levels; ++l) {
        if (l > 0) JC_ARG(ctx, ((nx >> (l - 1)) % 2 == 0) && ((ny >> (l - 1)) % 2 == 0) && ((nz >> (l - 1)) % 2 == 0));
        mg->dims[l] = Dim3i{(int)(nx >> l), (int)(ny >> l), (int)(nz >> l)};
        JC_ARG(ctx, A[l] != nullptr);
        DevMat* m = new DevMat(ctx);
        mats.push_back(m);
        JC_TRY(m->upload(A[l], mg->dims[l].n()));
        mg->A[l] = mat_of(m->A);
    }
    return mg_alloc_vectors(ctx, mg);
}

}  // namespace

#define JC_ENTER(ctx)                                    \
    if (!(ctx)) return JETCU_ERR_INVALID_ARG;            \
    if (cudaSetDevice((ctx)->device) != cudaSuccess)     \
        return jetcu_fail((ctx), JETCU_ERR_CUDA, "cudaSetDevice failed");

extern "C" {

// ---------------------------------------------------------------------------------------------- seam 1
int jetcu_fdm_mvm(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A, const double* v,
                  double* result) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && v && result);
    size_t n = nx * ny * nz;
    DevMat m(ctx); DevVec dv(ctx), out(ctx);
    JC_TRY(m.upload(A, n)); JC_TRY(dv.upload(v, n)); JC_TRY(out.upload(nullptr, n));
    JC_TRY(fdm_mvm(ctx, Dim3i{(int)nx, (int)ny, (int)nz}, mat_of(m.A), dv.p, out.p));
    JC_TRY(jc_d2h(ctx, result, out.p, n));
    return jc_sync(ctx);
}

int jetcu_fdm_residual(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A,
                       const double* x, const double* b, double* result) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && x && b && result);
    size_t n = nx * ny * nz;
    DevMat m(ctx); DevVec dx(ctx), db(ctx), out(ctx);
    JC_TRY(m.upload(A, n)); JC_TRY(dx.upload(x, n)); JC_TRY(db.upload(b, n)); JC_TRY(out.upload(nullptr, n));
    JC_TRY(fdm_residual(ctx, Dim3i{(int)nx, (int)ny, (int)nz}, mat_of(m.A), dx.p, db.p, out.p));
    JC_TRY(jc_d2h(ctx, result, out.p, n));
    return jc_sync(ctx);
}

int jetcu_fdm_relax_red_black(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A,
                              const double* b, double sor_factor, uint32_t iterations, double* x) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && x && b);
    size_t n = nx * ny * nz;
    DevMat m(ctx); DevVec dx(ctx), db(ctx), alt(ctx);
    JC_TRY(m.upload(A, n)); JC_TRY(dx.upload(x, n)); JC_TRY(db.upload(b, n)); JC_TRY(alt.upload(nullptr, n));
    JC_TRY(fdm_relax_rb(ctx, Dim3i{(int)nx, (int)ny, (int)nz}, mat_of(m.A), db.p, sor_factor, iterations, &dx.p, &alt.p));
    JC_TRY(jc_d2h(ctx, x, dx.p, n));
    return jc_sync(ctx);
}

int jetcu_fdm_relax(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A, const double* b,
                    double sor_factor, uint32_t iterations, double* x) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && x && b);
    size_t n = nx * ny * nz;
    DevMat m(ctx); DevVec dx(ctx), db(ctx);
    JC_TRY(m.upload(A, n)); JC_TRY(dx.upload(x, n)); JC_TRY(db.upload(b, n));
    JC_TRY(fdm_relax_lex(ctx, Dim3i{(int)nx, (int)ny, (int)nz}, mat_of(m.A), db.p, sor_factor, iterations, dx.p));
    JC_TRY(jc_d2h(ctx, x, dx.p, n));
    return jc_sync(ctx);
}

int jetcu_fdm_restrict(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const double* finer, double* coarser) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, finer && coarser);
    JC_ARG(ctx, nx % 2 == 0 && ny % 2 == 0 && nz % 2 == 0);  // JET_ASSERT(finer == 2*coarser)
    size_t n = nx * ny * nz;
    DevVec f(ctx), c(ctx);
    JC_TRY(f.upload(finer, n)); JC_TRY(c.upload(nullptr, n / 8));
    JC_TRY(fdm_restrict(ctx, Dim3i{(int)nx, (int)ny, (int)nz}, f.p, c.p));
    JC_TRY(jc_d2h(ctx, coarser, c.p, n / 8));
    return jc_sync(ctx);
}

int jetcu_fdm_correct(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const double* coarser, double* finer) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, finer && coarser);
    JC_ARG(ctx, nx % 2 == 0 && ny % 2 == 0 && nz % 2 == 0);
    size_t n = nx * ny * nz;
    DevVec f(ctx), c(ctx);
    JC_TRY(f.upload(finer, n)); JC_TRY(c.upload(coarser, n / 8));
    JC_TRY(fdm_correct(ctx, Dim3i{(int)nx, (int)ny, (int)nz}, c.p, f.p));
    JC_TRY(jc_d2h(ctx, finer, f.p, n));
    return jc_sync(ctx);
}

int jetcu_fdm_dot(jetcu_ctx* ctx, uint64_t n, const double* a, const double* b, double* result) {
    JC_ENTER(ctx);
    JC_ARG(ctx, a && b && result);
    DevVec da(ctx), db(ctx);
    JC_TRY(da.upload(a, n)); JC_TRY(db.upload(b, n));
    JC_TRY(fdm_dot(ctx, n, da.p, db.p, ctx->d_scalars));
    JC_TRY(jc_d2h(ctx, ctx->h_scalars, ctx->d_scalars, 1));
    JC_TRY(jc_sync(ctx));
    *result = ctx->h_scalars[0];
    return JETCU_OK;
}

int jetcu_fdm_axpy(jetcu_ctx* ctx, uint64_t n, double a, const double* x, const double* y, double* result) {
    JC_ENTER(ctx);
    JC_ARG(ctx, x && y && result);
    DevVec dx(ctx), dy(ctx), out(ctx);
    JC_TRY(dx.upload(x, n)); JC_TRY(dy.upload(y, n)); JC_TRY(out.upload(nullptr, n));
    JC_TRY(fdm_axpy(ctx, n, a, dx.p, dy.p, out.p));
    JC_TRY(jc_d2h(ctx, result, out.p, n));
    return jc_sync(ctx);
}

int jetcu_fdm_vcycle(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels,
                     const jetcu_matrix_row* const* A, const double* b, double* x, const jetcu_mgpcg_params* p) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && b && x && p);
    MgLevels mg;
    std::vector<DevMat*> mats;
    int rc = setup_levels(ctx, &mg, mats, nx, ny, nz, levels, A);
    size_t n = nx * ny * nz;
    if (rc == JETCU_OK) rc = jc_h2d(ctx, mg.x[0], x, n);
    if (rc == JETCU_OK) rc = jc_h2d(ctx, mg.b[0], b, n);
    if (rc == JETCU_OK) rc = mg_vcycle(ctx, &mg, p, 0);
    if (rc == JETCU_OK) rc = gather_level0(ctx, &mg, mg.x[0]);
    if (rc == JETCU_OK) rc = jc_d2h(ctx, x, mg.x[0], n);
    if (rc == JETCU_OK) rc = jc_sync(ctx);
    mg_free(ctx, &mg);
    for (auto* m : mats) delete m;
    return rc;
}

int jetcu_mgpcg_solve(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels,
                      const jetcu_matrix_row* const* A, const double* b, double* x, const jetcu_mgpcg_params* p,
                      jetcu_mgpcg_result* result) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && b && x && p);
    MgLevels mg;
    std::vector<DevMat*> mats;
    size_t n = nx * ny * nz;
    DevVec db(ctx), dx(ctx);
    int rc = setup_levels(ctx, &mg, mats, nx, ny, nz, levels, A);
    if (rc == JETCU_OK) rc = db.upload(b, n);
    if (rc == JETCU_OK) rc = dx.upload(nullptr, n);
    if (rc == JETCU_OK) rc = mgpcg_solve_device(ctx, &mg, db.p, dx.p, p, result);
    if (rc == JETCU_OK) rc = jc_d2h(ctx, x, dx.p, n);
    if (rc == JETCU_OK) rc = jc_sync(ctx);
    mg_free(ctx, &mg);
    for (auto* m : mats) delete m;
    return rc;
}

int jetcu_iccg_solve(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A, const double* b, double* x,
                     uint32_t max_iterations, double tolerance, jetcu_mgpcg_result* result) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && b && x);
    const size_t n = nx * ny * nz;
    DevMat m(ctx);
    DevVec db(ctx), dx(ctx);
    JC_TRY(m.upload(A, n)); JC_TRY(db.upload(b, n)); JC_TRY(dx.upload(nullptr, n));
    JC_TRY(iccg_solve_device(ctx, Dim3i{(int)nx, (int)ny, (int)nz}, mat_of(m.A), db.p, dx.p, max_iterations, tolerance, result));
    JC_TRY(jc_d2h(ctx, x, dx.p, n));
    return jc_sync(ctx);
}

int jetcu_iccg_precondition(jetcu_ctx* ctx, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A, const double* b,
                            double* d, double* y, double* x) {
    JC_ENTER(ctx);
    JC_TRY(check_dims(ctx, nx, ny, nz));
    JC_ARG(ctx, A && b && y && x);
    const size_t n = nx * ny * nz;
    const Dim3i dim{(int)nx, (int)ny, (int)nz};
    DevMat m(ctx);
    DevVec db(ctx), dd(ctx), dy(ctx), dx(ctx);
    JC_TRY(m.upload(A, n)); JC_TRY(db.upload(b, n)); JC_TRY(dd.upload(nullptr, n)); JC_TRY(dy.upload(nullptr, n));
    JC_TRY(dx.upload(nullptr, n));
    JC_TRY(iccg_build(ctx, dim, mat_of(m.A), dd.p));
    JC_TRY(iccg_precondition(ctx, dim, mat_of(m.A), dd.p, db.p, dy.p, dx.p));
    if (d) JC_TRY(jc_d2h(ctx, d, dd.p, n));
    JC_TRY(jc_d2h(ctx, y, dy.p, n)); JC_TRY(jc_d2h(ctx, x, dx.p, n));
    return jc_sync(ctx);
}

// ---------------------------------------------------------------------------------------------- seam 2
int jetcu_pressure_solve(jetcu_ctx* ctx, const jetcu_grid* g, const double* u_in, const double* v_in,
                         const double* w_in, const double* collider_sdf, const double* bnd_u, const double* bnd_v,
                         const double* bnd_w, const double* fluid_sdf, const jetcu_mgpcg_params* p, double* u_out,
                         double* v_out, double* w_out, double* pressure_out, jetcu_mgpcg_result* result) {
    JC_ENTER(ctx);
    JC_ARG(ctx, g && p && u_in && v_in && w_in && u_out && v_out && w_out);
    JC_TRY(check_dims(ctx, g->nx, g->ny, g->nz));
    JC_ARG(ctx, g->hx > 0 && g->hy > 0 && g->hz > 0);
    uint32_t levels = jetcu_mg_num_levels(g->nx, g->ny, g->nz, p->max_levels);
    size_t nc = g->nx * g->ny * g->nz;
    size_t nu = (g->nx + 1) * g->ny * g->nz, nv = g->nx * (g->ny + 1) * g->nz, nw = g->nx * g->ny * (g->nz + 1);
    JC_TRY(pressure_prepare(ctx, ctx->psys, g, p->max_levels));
    DevVec du(ctx), dv(ctx), dw(ctx), duo(ctx), dvo(ctx), dwo(ctx), dc(ctx), df(ctx), dbu(ctx), dbv(ctx), dbw(ctx);
    JC_TRY(du.upload(u_in, nu)); JC_TRY(dv.upload(v_in, nv)); JC_TRY(dw.upload(w_in, nw));
    JC_TRY(duo.upload(u_out, nu)); JC_TRY(dvo.upload(v_out, nv)); JC_TRY(dwo.upload(w_out, nw));
    if (collider_sdf) JC_TRY(dc.upload(collider_sdf, nc));
    if (fluid_sdf) JC_TRY(df.upload(fluid_sdf, nc));
    if (bnd_u) JC_TRY(dbu.upload(bnd_u, nu));
    if (bnd_v) JC_TRY(dbv.upload(bnd_v, nv));
    if (bnd_w) JC_TRY(dbw.upload(bnd_w, nw));
    JC_TRY(pressure_solve_device(ctx, ctx->psys, du.p, dv.p, dw.p, dc.p, dbu.p, dbv.p, dbw.p, df.p, p, duo.p, dvo.p,
                                 dwo.p, result));
    JC_TRY(jc_d2h(ctx, u_out, duo.p, nu)); JC_TRY(jc_d2h(ctx, v_out, dvo.p, nv)); JC_TRY(jc_d2h(ctx, w_out, dwo.p, nw));
    if (pressure_out) JC_TRY(jc_d2h(ctx, pressure_out, ctx->psys->x0, nc));
    return jc_sync(ctx);
}

int jetcu_pressure_get_system(jetcu_ctx* ctx, uint32_t level, jetcu_matrix_row* A, double* b, float* fluid_sdf,
                              float* uw, float* vw, float* ww) {
    JC_ENTER(ctx);
    PressureSystem* ps = ctx->psys;
    JC_ARG(ctx, ps && (int)level < ps->levels);
    Dim3i d = ps->dims[level];
    size_t n = d.n();
    if (A) {
        jetcu_matrix_row* aos = nullptr;
        JC_TRY(jc_alloc(ctx, &aos, n));
        int rc = fdm_mat_to_aos(ctx, ps->mg.A[level], aos, n);
        if (rc == JETCU_OK) rc = jc_d2h(ctx, A, aos, n);
        if (rc == JETCU_OK) rc = jc_sync(ctx);
        jc_free(ctx, aos);
        JC_TRY(rc);
    }
    if (b) {
        if (level == 0) JC_TRY(jc_d2h(ctx, b, ps->b0, n)); else memset(b, 0, n * sizeof(double));
    }
    if (fluid_sdf) JC_TRY(jc_d2h(ctx, fluid_sdf, ps->sdf[level], n));
    if (uw) JC_TRY(jc_d2h(ctx, uw, ps->uw[level], (size_t)(d.nx + 1) * d.ny * d.nz));
    if (vw) JC_TRY(jc_d2h(ctx, vw, ps->vw[level], (size_t)d.nx * (d.ny + 1) * d.nz));
    if (ww) JC_TRY(jc_d2h(ctx, ww, ps->ww[level], (size_t)d.nx * d.ny * (d.nz + 1)));
    return jc_sync(ctx);
}

int jetcu_pressure_vcycle(jetcu_ctx* ctx, const double* b, double* x, const jetcu_mgpcg_params* p) {
    JC_ENTER(ctx);
    PressureSystem* ps = ctx->psys;
    JC_ARG(ctx, ps && ps->levels > 0 && b && x && p);
    MgLevels* mg = &ps->mg;
    size_t n = ps->dims[0].n();
    JC_TRY(jc_h2d(ctx, mg->x[0], x, n));
    JC_TRY(jc_h2d(ctx, mg->b[0], b, n));
    JC_TRY(mg_make_plan(ctx, mg));
    JC_TRY(mg_vcycle(ctx, mg, p, 0));
    JC_TRY(gather_level0(ctx, mg, mg->x[0]));
    JC_TRY(jc_d2h(ctx, x, mg->x[0], n));
    return jc_sync(ctx);
}

}  // extern "C"
