// ref_shim.cpp — TEST INFRASTRUCTURE.  extern "C" view (oracle/jet_oracle.h) of the UNMODIFIED Jet
// reference, compiled from /root/reference by oracle/Makefile into oracle/_ref/libjetref.so.
// Nothing from the reference is copied here: this file only *calls* Jet's public classes.
// `private`/`protected` are opened for this one translation unit so the shim can read the solver's
// intermediate state (weights, matrices, markers) and drive individual stages for stage-wise parity.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <fstream>
#include <sstream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <vector>
#include <omp.h>

#define private public
#define protected public
#include <jet/jet.h>
#undef private
#undef protected

#include "jet_oracle.h"

using namespace jet;

namespace {

struct MuteOnce {
    MuteOnce() { Logging::mute(); }
} g_muteOnce;

double nowMs() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void fillMatrix(FdmMatrix3* dst, const jeto_row* src, Size3 size) {
    dst->resize(size);
    static_assert(sizeof(FdmMatrixRow3) == sizeof(jeto_row), "row layout");
    std::memcpy(dst->data(), src, sizeof(jeto_row) * size.x * size.y * size.z);
}

void fillVector(FdmVector3* dst, const double* src, Size3 size) {
    dst->resize(size);
    if (src) std::memcpy(dst->data(), src, sizeof(double) * size.x * size.y * size.z);
}

void copyOut(double* dst, const FdmVector3& src) {
    std::memcpy(dst, src.data(), sizeof(double) * src.size().x * src.size().y * src.size().z);
}

std::shared_ptr<FdmMgpcgSolver3> makeMgpcg(const jeto_mgpcg_params* p) {
    return std::make_shared<FdmMgpcgSolver3>(p->max_iterations, p->max_levels, p->n_restriction, p->n_correction,
                                             p->n_coarsest, p->n_final, p->tolerance, p->sor_factor,
                                             p->use_red_black != 0);
}

void buildMgSystem(FdmMgLinearSystem3* sys, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels,
                   const jeto_row* const* A, const double* b, const double* x) {
    sys->resizeWithFinest(Size3(nx, ny, nz), levels);
    for (uint32_t l = 0; l < sys->A.levels.size(); ++l) {
        fillMatrix(&sys->A.levels[l], A[l], sys->A.levels[l].size());
    }
    fillVector(&sys->b.levels.front(), b, Size3(nx, ny, nz));
    if (x) fillVector(&sys->x.levels.front(), x, Size3(nx, ny, nz));
}

double trueResidual(const FdmMatrix3& A, const FdmVector3& x, const FdmVector3& b) {
    FdmVector3 r(A.size());
    FdmBlas3::residual(A, x, b, &r);
    return FdmBlas3::l2Norm(r);
}

std::unique_ptr<GridFractionalSinglePhasePressureSolver3> g_lastPressureSolver;

}  // namespace

extern "C" {

const char* jeto_kind(void) { return "reference"; }
void jeto_set_num_threads(int n) {
    if (n <= 0) n = omp_get_num_procs();
    setMaxNumberOfThreads(static_cast<unsigned int>(n));
    omp_set_num_threads(n);
}
int jeto_num_threads(void) { return static_cast<int>(maxNumberOfThreads()); }

int jeto_fdm_mvm(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* v, double* result) {
    Size3 s(nx, ny, nz);
    FdmMatrix3 m; FdmVector3 vv, r(s);
    fillMatrix(&m, A, s); fillVector(&vv, v, s);
    FdmBlas3::mvm(m, vv, &r);
    copyOut(result, r);
    return 0;
}

int jeto_fdm_residual(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* x, const double* b,
                      double* result) {
    Size3 s(nx, ny, nz);
    FdmMatrix3 m; FdmVector3 xx, bb, r(s);
    fillMatrix(&m, A, s); fillVector(&xx, x, s); fillVector(&bb, b, s);
    FdmBlas3::residual(m, xx, bb, &r);
    copyOut(result, r);
    return 0;
}

int jeto_fdm_relax_red_black(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b,
                             double sor_factor, uint32_t iterations, double* x) {
    Size3 s(nx, ny, nz);
    FdmMatrix3 m; FdmVector3 xx, bb;
    fillMatrix(&m, A, s); fillVector(&xx, x, s); fillVector(&bb, b, s);
    for (uint32_t it = 0; it < iterations; ++it) FdmGaussSeidelSolver3::relaxRedBlack(m, bb, sor_factor, &xx);
    copyOut(x, xx);
    return 0;
}

int jeto_fdm_relax(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b, double sor_factor,
                   uint32_t iterations, double* x) {
    Size3 s(nx, ny, nz);
    FdmMatrix3 m; FdmVector3 xx, bb;
    fillMatrix(&m, A, s); fillVector(&xx, x, s); fillVector(&bb, b, s);
    for (uint32_t it = 0; it < iterations; ++it) FdmGaussSeidelSolver3::relax(m, bb, sor_factor, &xx);
    copyOut(x, xx);
    return 0;
}

int jeto_fdm_restrict(uint64_t nx, uint64_t ny, uint64_t nz, const double* finer, double* coarser) {
    FdmVector3 f, c(Size3(nx / 2, ny / 2, nz / 2));
    fillVector(&f, finer, Size3(nx, ny, nz));
    FdmMgUtils3::restrict(f, &c);
    copyOut(coarser, c);
    return 0;
}

int jeto_fdm_correct(uint64_t nx, uint64_t ny, uint64_t nz, const double* coarser, double* finer) {
    if (ny != nz) return 1;  // reference reads out of bounds otherwise (fdm_mg_linear_system3.cpp:159)
    FdmVector3 f, c;
    fillVector(&f, finer, Size3(nx, ny, nz));
    fillVector(&c, coarser, Size3(nx / 2, ny / 2, nz / 2));
    FdmMgUtils3::correct(c, &f);
    copyOut(finer, f);
    return 0;
}

int jeto_fdm_dot(uint64_t n, const double* a, const double* b, double* result) {
    FdmVector3 aa, bb;
    fillVector(&aa, a, Size3(n, 1, 1)); fillVector(&bb, b, Size3(n, 1, 1));
    *result = FdmBlas3::dot(aa, bb);
    return 0;
}

int jeto_fdm_axpy(uint64_t n, double a, const double* x, const double* y, double* result) {
    FdmVector3 xx, yy, rr(Size3(n, 1, 1));
    fillVector(&xx, x, Size3(n, 1, 1)); fillVector(&yy, y, Size3(n, 1, 1));
    FdmBlas3::axpy(a, xx, yy, &rr);
    copyOut(result, rr);
    return 0;
}

uint32_t jeto_mg_num_levels(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t max_levels) {
    std::vector<Array3<float>> levels;
    FdmMgUtils3::resizeArrayWithFinest(Size3(nx, ny, nz), max_levels, &levels);
    return static_cast<uint32_t>(levels.size());
}

int jeto_fdm_vcycle(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A,
                    const double* b, double* x, const jeto_mgpcg_params* p) {
    FdmMgLinearSystem3 sys;
    buildMgSystem(&sys, nx, ny, nz, levels, A, b, x);
    FdmMgSolver3 solver(p->max_levels, p->n_restriction, p->n_correction, p->n_coarsest, p->n_final, p->tolerance,
                        p->sor_factor, p->use_red_black != 0);
    solver.solve(&sys);
    copyOut(x, sys.x.levels.front());
    return 0;
}

int jeto_mgpcg_solve(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A,
                     const double* b, double* x, const jeto_mgpcg_params* p, jeto_mgpcg_result* result) {
    FdmMgLinearSystem3 sys;
    buildMgSystem(&sys, nx, ny, nz, levels, A, b, nullptr);
    auto solver = makeMgpcg(p);
    double t0 = nowMs();
    bool ret = solver->solve(&sys);
    double t1 = nowMs();
    copyOut(x, sys.x.levels.front());
    if (result) {
        result->iterations = solver->lastNumberOfIterations();
        result->returned = ret ? 1 : 0;
        result->last_residual = solver->lastResidual();
        result->true_residual_l2 = trueResidual(sys.A.levels.front(), sys.x.levels.front(), sys.b.levels.front());
        result->b_l2 = FdmBlas3::l2Norm(sys.b.levels.front());
        result->solve_ms = t1 - t0;
    }
    return 0;
}

// FdmIccgSolver3::solve(FdmLinearSystem3*) — the reference's own object on a caller-supplied system
int jeto_iccg_solve(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b, double* x,
                    uint32_t max_iterations, double tolerance, jeto_mgpcg_result* result) {
    FdmLinearSystem3 sys;
    const Size3 size(nx, ny, nz);
    sys.A.resize(size); sys.b.resize(size); sys.x.resize(size);
    fillMatrix(&sys.A, A, size);
    fillVector(&sys.b, b, size);
    FdmIccgSolver3 solver(max_iterations, tolerance);
    double t0 = nowMs();
    bool ret = solver.solve(&sys);
    double t1 = nowMs();
    copyOut(x, sys.x);
    if (result) {
        result->iterations = solver.lastNumberOfIterations();
        result->returned = ret ? 1 : 0;
        result->last_residual = solver.lastResidual();
        result->true_residual_l2 = trueResidual(sys.A, sys.x, sys.b);
        result->b_l2 = FdmBlas3::l2Norm(sys.b);
        result->solve_ms = t1 - t0;
    }
    return 0;
}

// FdmIccgSolver3::Preconditioner::build / solve on their own
int jeto_iccg_precondition(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b, double* d, double* y,
                           double* x) {
    const Size3 size(nx, ny, nz);
    FdmMatrix3 m(size);
    FdmVector3 rhs(size), out(size);
    fillMatrix(&m, A, size);
    fillVector(&rhs, b, size);
    FdmIccgSolver3::Preconditioner pre;
    pre.build(m);
    pre.solve(rhs, &out);
    if (d) copyOut(d, pre.d);
    copyOut(y, pre.y);
    copyOut(x, out);
    return 0;
}

int jeto_pressure_solve(const jeto_grid* g, const double* u_in, const double* v_in, const double* w_in,
                        const double* collider_sdf, const double* bnd_u, const double* bnd_v, const double* bnd_w,
                        const double* fluid_sdf, const jeto_mgpcg_params* p, double* u_out, double* v_out,
                        double* w_out, double* pressure_out, jeto_mgpcg_result* result) {
    Size3 res(g->nx, g->ny, g->nz);
    Vector3D h(g->hx, g->hy, g->hz), o(g->ox, g->oy, g->oz);
    FaceCenteredGrid3 input(res, h, o);
    FaceCenteredGrid3 output(res, h, o);
    size_t nu = (g->nx + 1) * g->ny * g->nz, nv = g->nx * (g->ny + 1) * g->nz, nw = g->nx * g->ny * (g->nz + 1);
    std::memcpy(input.uAccessor().data(), u_in, nu * sizeof(double));
    std::memcpy(input.vAccessor().data(), v_in, nv * sizeof(double));
    std::memcpy(input.wAccessor().data(), w_in, nw * sizeof(double));
    std::memcpy(output.uAccessor().data(), u_out, nu * sizeof(double));
    std::memcpy(output.vAccessor().data(), v_out, nv * sizeof(double));
    std::memcpy(output.wAccessor().data(), w_out, nw * sizeof(double));

    ScalarField3Ptr bsdf, fsdf;
    if (collider_sdf) {
        auto grid = std::make_shared<CellCenteredScalarGrid3>(res, h, o);
        std::memcpy(grid->dataAccessor().data(), collider_sdf, sizeof(double) * g->nx * g->ny * g->nz);
        bsdf = grid;
    } else {
        bsdf = std::make_shared<ConstantScalarField3>(kMaxD);
    }
    if (fluid_sdf) {
        auto grid = std::make_shared<CellCenteredScalarGrid3>(res, h, o);
        std::memcpy(grid->dataAccessor().data(), fluid_sdf, sizeof(double) * g->nx * g->ny * g->nz);
        fsdf = grid;
    } else {
        fsdf = std::make_shared<ConstantScalarField3>(-kMaxD);
    }
    VectorField3Ptr bvel;
    if (bnd_u || bnd_v || bnd_w) {
        // boundaryVelocity is only ever sampled at face positions, one component per face family
        // (grid_fractional_single_phase_pressure_solver3.cpp:242-255); decode the face index by rounding.
        Vector3D ou = input.uOrigin(), ov = input.vOrigin(), ow = input.wOrigin();
        auto fn = [=](const Vector3D& x) {
            Vector3D out;
            auto idx = [&](const Vector3D& org, size_t sx, size_t sy, size_t sz, const double* arr) -> double {
                if (!arr) return 0.0;
                long i = std::lround((x.x - org.x) / h.x), j = std::lround((x.y - org.y) / h.y),
                     k = std::lround((x.z - org.z) / h.z);
                i = std::min<long>(std::max<long>(i, 0), sx - 1);
                j = std::min<long>(std::max<long>(j, 0), sy - 1);
                k = std::min<long>(std::max<long>(k, 0), sz - 1);
                return arr[(k * sy + j) * sx + i];
            };
            out.x = idx(ou, res.x + 1, res.y, res.z, bnd_u);
            out.y = idx(ov, res.x, res.y + 1, res.z, bnd_v);
            out.z = idx(ow, res.x, res.y, res.z + 1, bnd_w);
            return out;
        };
        bvel = std::make_shared<CustomVectorField3>(fn);
    } else {
        bvel = std::make_shared<ConstantVectorField3>(Vector3D());
    }

    g_lastPressureSolver.reset(new GridFractionalSinglePhasePressureSolver3());
    if (p->linear_solver & 1) {   // FdmIccgSolver3, optionally on the compressed (liquid cells only) system
        auto iccg = std::make_shared<FdmIccgSolver3>(p->max_iterations, p->tolerance);
        g_lastPressureSolver->setLinearSystemSolver(iccg);
        const bool compressed = (p->linear_solver & 2) != 0;
        double t0 = nowMs();
        g_lastPressureSolver->solve(input, 0.0, &output, *bsdf, *bvel, *fsdf, compressed);
        double t1 = nowMs();
        std::memcpy(u_out, output.uAccessor().data(), nu * sizeof(double));
        std::memcpy(v_out, output.vAccessor().data(), nv * sizeof(double));
        std::memcpy(w_out, output.wAccessor().data(), nw * sizeof(double));
        if (pressure_out) copyOut(pressure_out, g_lastPressureSolver->pressure());
        if (result) {
            result->iterations = iccg->lastNumberOfIterations();
            result->last_residual = iccg->lastResidual();
            result->returned = (iccg->lastResidual() <= iccg->tolerance() ||
                                iccg->lastNumberOfIterations() < iccg->maxNumberOfIterations()) ? 1 : 0;
            if (compressed) {
                const auto& cs = g_lastPressureSolver->_compSystem;
                VectorND r(cs.b.size());
                FdmCompressedBlas3::residual(cs.A, cs.x, cs.b, &r);
                result->true_residual_l2 = FdmCompressedBlas3::l2Norm(r);
                result->b_l2 = FdmCompressedBlas3::l2Norm(cs.b);
            } else {
                const auto& sys = g_lastPressureSolver->_system;
                result->true_residual_l2 = trueResidual(sys.A, sys.x, sys.b);
                result->b_l2 = FdmBlas3::l2Norm(sys.b);
            }
            result->solve_ms = t1 - t0;
        }
        return 0;
    }
    auto lin = makeMgpcg(p);
    g_lastPressureSolver->setLinearSystemSolver(lin);
    double t0 = nowMs();
    g_lastPressureSolver->solve(input, 0.0, &output, *bsdf, *bvel, *fsdf, false);
    double t1 = nowMs();

    std::memcpy(u_out, output.uAccessor().data(), nu * sizeof(double));
    std::memcpy(v_out, output.vAccessor().data(), nv * sizeof(double));
    std::memcpy(w_out, output.wAccessor().data(), nw * sizeof(double));
    if (pressure_out) copyOut(pressure_out, g_lastPressureSolver->pressure());
    if (result) {
        const auto& sys = g_lastPressureSolver->_mgSystem;
        result->iterations = lin->lastNumberOfIterations();
        result->last_residual = lin->lastResidual();
        result->returned = (lin->lastResidual() <= lin->tolerance() ||
                            lin->lastNumberOfIterations() < lin->maxNumberOfIterations()) ? 1 : 0;
        result->true_residual_l2 = trueResidual(sys.A.levels.front(), sys.x.levels.front(), sys.b.levels.front());
        result->b_l2 = FdmBlas3::l2Norm(sys.b.levels.front());
        result->solve_ms = t1 - t0;
    }
    return 0;
}

int jeto_pressure_get_system(uint32_t level, jeto_row* A, double* b, float* fluid_sdf, float* u_weights,
                             float* v_weights, float* w_weights) {
    if (!g_lastPressureSolver) return 1;
    auto& s = *g_lastPressureSolver;
    if (level >= s._mgSystem.A.levels.size()) return 1;
    auto n3 = [](const Size3& sz) { return sz.x * sz.y * sz.z; };
    if (A) std::memcpy(A, s._mgSystem.A.levels[level].data(), sizeof(jeto_row) * n3(s._mgSystem.A.levels[level].size()));
    if (b) std::memcpy(b, s._mgSystem.b.levels[level].data(), sizeof(double) * n3(s._mgSystem.b.levels[level].size()));
    if (fluid_sdf) std::memcpy(fluid_sdf, s._fluidSdf[level].data(), sizeof(float) * n3(s._fluidSdf[level].size()));
    if (u_weights) std::memcpy(u_weights, s._uWeights[level].data(), sizeof(float) * n3(s._uWeights[level].size()));
    if (v_weights) std::memcpy(v_weights, s._vWeights[level].data(), sizeof(float) * n3(s._vWeights[level].size()));
    if (w_weights) std::memcpy(w_weights, s._wWeights[level].data(), sizeof(float) * n3(s._wWeights[level].size()));
    return 0;
}

int jeto_bin_points(uint64_t n, const double* positions, uint64_t rx, uint64_t ry, uint64_t rz, double spacing,
                    uint64_t* keys_unsorted, uint64_t* sorted_indices, uint64_t* keys_sorted, uint64_t* start_table,
                    uint64_t* end_table) {
    PointParallelHashGridSearcher3 searcher(Size3(rx, ry, rz), spacing);
    Array1<Vector3D> pts(n);
    for (uint64_t i = 0; i < n; ++i) pts[i] = Vector3D(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
    searcher.build(pts.constAccessor());
    if (keys_unsorted)
        for (uint64_t i = 0; i < n; ++i) keys_unsorted[i] = searcher.getHashKeyFromPosition(pts[i]);
    static_assert(sizeof(size_t) == sizeof(uint64_t), "size_t");
    if (sorted_indices) std::memcpy(sorted_indices, searcher.sortedIndices().data(), n * 8);
    if (keys_sorted) std::memcpy(keys_sorted, searcher.keys().data(), n * 8);
    if (start_table) std::memcpy(start_table, searcher.startIndexTable().data(), rx * ry * rz * 8);
    if (end_table) std::memcpy(end_table, searcher.endIndexTable().data(), rx * ry * rz * 8);
    return 0;
}

int jeto_bin_points_serial(uint64_t n, const double* positions, uint64_t rx, uint64_t ry, uint64_t rz, double spacing,
                           uint64_t* bucket_of_point, uint64_t* rank_in_bucket) {
    PointHashGridSearcher3 searcher(Size3(rx, ry, rz), spacing);
    Array1<Vector3D> pts(n);
    for (uint64_t i = 0; i < n; ++i) pts[i] = Vector3D(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
    searcher.build(pts.constAccessor());
    const auto& buckets = searcher.buckets();
    for (size_t key = 0; key < buckets.size(); ++key) {
        for (size_t r = 0; r < buckets[key].size(); ++r) {
            bucket_of_point[buckets[key][r]] = key;
            rank_in_bucket[buckets[key][r]] = r;
        }
    }
    return 0;
}

// ------------------------------------------------------------------ hybrid solver
struct jeto_hybrid {
    std::shared_ptr<PicSolver3> solver;
    std::shared_ptr<FdmMgpcgSolver3> lin;
    std::shared_ptr<FdmIccgSolver3> iccg;   // when params.mgpcg.linear_solver & 1
    std::shared_ptr<GridFractionalSinglePhasePressureSolver3> pressure;
    int mode;
};

int jeto_hybrid_create(const jeto_hybrid_params* p, jeto_hybrid** out) {
    auto* h = new jeto_hybrid();
    Size3 res(p->grid.nx, p->grid.ny, p->grid.nz);
    Vector3D hh(p->grid.hx, p->grid.hy, p->grid.hz), o(p->grid.ox, p->grid.oy, p->grid.oz);
    h->mode = p->mode;
    if (p->mode == 0) {
        h->solver = std::make_shared<PicSolver3>(res, hh, o);
    } else if (p->mode == 1) {
        auto s = std::make_shared<FlipSolver3>(res, hh, o);
        s->setPicBlendingFactor(p->pic_blending_factor);
        h->solver = s;
    } else {
        h->solver = std::make_shared<ApicSolver3>(res, hh, o);
    }
    h->pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
    if (p->mgpcg.linear_solver & 1) {   // Jet's default solver, as the hybrid_liquid_sim examples run it
        h->iccg = std::make_shared<FdmIccgSolver3>(p->mgpcg.max_iterations, p->mgpcg.tolerance);
        h->pressure->setLinearSystemSolver(h->iccg);
    } else {
        h->lin = makeMgpcg(&p->mgpcg);
        h->pressure->setLinearSystemSolver(h->lin);
    }
    h->solver->setPressureSolver(h->pressure);
    h->solver->setGravity(Vector3D(p->gravity[0], p->gravity[1], p->gravity[2]));
    h->solver->setMaxCfl(p->max_cfl);
    h->solver->setClosedDomainBoundaryFlag(p->closed_domain_boundary_flag);
    h->solver->setUseCompressedLinearSystem((p->mgpcg.linear_solver & 2) != 0);
    *out = h;
    return 0;
}

int jeto_hybrid_destroy(jeto_hybrid* h) { delete h; return 0; }

static Surface3Ptr makeShape(const jeto_collider_shape& s) {
    const Vector3D a(s.a[0], s.a[1], s.a[2]), b(s.b[0], s.b[1], s.b[2]);
    const bool flipped = s.flipped != 0;
    switch (s.kind) {
        case 1: return std::make_shared<Box3>(a, b, Transform3(), flipped);
        case 2: return std::make_shared<Sphere3>(a, s.b[0], Transform3(), flipped);
        case 3: return std::make_shared<Cylinder3>(a, s.b[0], s.b[1], Transform3(), flipped);
        case 4: return std::make_shared<Plane3>(b, a, Transform3(), flipped);   // (normal, point)
        default: return nullptr;
    }
}

int jeto_hybrid_set_collider_shapes(jeto_hybrid* h, const jeto_collider_shape* shapes, uint32_t n, int as_set, const double lin[3],
                                    const double ang[3], double friction) {
    if (n == 0) { h->solver->setCollider(nullptr); return 0; }
    if (n != 1 && !as_set) return 1;
    Surface3Ptr surface;
    if (as_set) {
        std::vector<Surface3Ptr> members;
        for (uint32_t i = 0; i < n; ++i) {
            auto m = makeShape(shapes[i]);
            if (!m) return 1;
            members.push_back(m);
        }
        surface = std::make_shared<ImplicitSurfaceSet3>(members);   // addExplicitSurface each: SurfaceToImplicit3 wrappers
    } else {
        surface = makeShape(shapes[0]);
        if (!surface) return 1;
    }
    auto col = std::make_shared<RigidBodyCollider3>(surface,
                                                    lin ? Vector3D(lin[0], lin[1], lin[2]) : Vector3D(),
                                                    ang ? Vector3D(ang[0], ang[1], ang[2]) : Vector3D());
    col->setFrictionCoefficient(friction);
    h->solver->setCollider(col);
    // what GridFluidSolver3::beginAdvanceTimeStep does before anything reads the collider grid (:448-453)
    auto grids = h->solver->gridSystemData();
    h->solver->_boundaryConditionSolver->updateCollider(col, grids->resolution(), grids->gridSpacing(), grids->origin());
    return 0;
}

int jeto_hybrid_set_collider_shape(jeto_hybrid* h, int kind, const double a[3], const double b[3], const double lin[3],
                                   const double ang[3], double friction) {
    if (kind < 0 || kind > 2) return 1;
    if (kind == 0) return jeto_hybrid_set_collider_shapes(h, nullptr, 0, 0, lin, ang, friction);
    jeto_collider_shape s{};
    s.kind = kind;
    for (int q = 0; q < 3; ++q) { s.a[q] = a[q]; s.b[q] = b ? b[q] : 0.0; }
    return jeto_hybrid_set_collider_shapes(h, &s, 1, 0, lin, ang, friction);
}

int jeto_hybrid_emit_shapes(jeto_hybrid* h, const jeto_collider_shape* shapes, uint32_t n, int as_set, int lattice, double spacing,
                            const double v0[3], uint64_t* n_emitted) {
    if (n == 0 || (n != 1 && !as_set)) return 1;
    auto ps = h->solver->particleSystemData();
    size_t before = ps->numberOfParticles();
    Surface3Ptr surface;
    if (as_set) {
        std::vector<Surface3Ptr> members;
        for (uint32_t i = 0; i < n; ++i) {
            auto m = makeShape(shapes[i]);
            if (!m) return 1;
            members.push_back(m);
        }
        surface = ImplicitSurfaceSet3::builder().withExplicitSurfaces(members).makeShared();   // as main.cpp:236-238
    } else {
        surface = makeShape(shapes[0]);
        if (!surface) return 1;
    }
    auto emitter = VolumeParticleEmitter3::builder()
                       .withSurface(surface)
                       .withMaxRegion(h->solver->gridSystemData()->boundingBox())
                       .withSpacing(spacing)
                       .withInitialVelocity(v0 ? Vector3D(v0[0], v0[1], v0[2]) : Vector3D())
                       .withIsOneShot(true)
                       .withJitter(0.0)
                       .makeShared();
    if (lattice == 0) emitter->setPointGenerator(std::make_shared<GridPointGenerator3>());   // else the default: BccLatticePointGenerator
    emitter->setTarget(ps);
    emitter->update(0.0, 0.0);
    if (n_emitted) *n_emitted = ps->numberOfParticles() - before;
    return 0;
}

// jet::serialize(const uint8_t*, size_t, std::vector<uint8_t>*) — src/jet/serialization.cpp:25-38 (the framing of the
// examples' .pos files: saveParticleAsPos, src/examples/hybrid_liquid_sim/main.cpp:29-45)
int jeto_serialize_bytes(const uint8_t* data, uint64_t size, uint8_t* out, uint64_t cap, uint64_t* out_size) {
    std::vector<uint8_t> buffer;
    jet::serialize(data, (size_t)size, &buffer);
    *out_size = buffer.size();
    if (buffer.size() > cap) return 1;
    std::memcpy(out, buffer.data(), buffer.size());
    return 0;
}

// what saveParticleAsXyz streams per particle (main.cpp:47-63): operator<< on doubles with the default format
int jeto_format_xyz(const double* pos, uint64_t n, char* out, uint64_t cap, uint64_t* out_size) {
    std::ostringstream os;
    for (uint64_t i = 0; i < n; ++i) os << pos[3 * i] << ' ' << pos[3 * i + 1] << ' ' << pos[3 * i + 2] << std::endl;
    const std::string s = os.str();
    *out_size = s.size();
    if (s.size() > cap) return 1;
    std::memcpy(out, s.data(), s.size());
    return 0;
}

int jeto_hybrid_get_collider_sdf(jeto_hybrid* h, double* sdf) {
    auto field = h->solver->colliderSdf();
    auto grid = std::dynamic_pointer_cast<CellCenteredScalarGrid3>(field);
    auto res = h->solver->gridSystemData()->resolution();
    size_t n = res.x * res.y * res.z;
    if (!grid) { for (size_t i = 0; i < n; ++i) sdf[i] = kMaxD; return 0; }
    std::memcpy(sdf, grid->constDataAccessor().data(), n * sizeof(double));
    return 0;
}

int jeto_hybrid_set_particles(jeto_hybrid* h, uint64_t n, const double* pos, const double* vel) {
    auto ps = h->solver->particleSystemData();
    ps->resize(n);
    auto P = ps->positions(); auto V = ps->velocities();
    for (uint64_t i = 0; i < n; ++i) {
        P[i] = Vector3D(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        V[i] = vel ? Vector3D(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]) : Vector3D();
    }
    return 0;
}

// ParticleSystemData3::addParticles (src/jet/particle_system_data3.cpp:158-197): what a host emitter does between steps
int jeto_hybrid_append_particles(jeto_hybrid* h, uint64_t n, const double* pos, const double* vel) {
    Array1<Vector3D> P(n), V(n);
    for (uint64_t i = 0; i < n; ++i) {
        P[i] = Vector3D(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        V[i] = vel ? Vector3D(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]) : Vector3D();
    }
    h->solver->particleSystemData()->addParticles(P.constAccessor(), V.constAccessor());
    return 0;
}

int jeto_hybrid_emit_box(jeto_hybrid* h, const double lower[3], const double upper[3], double spacing,
                         const double v0[3], uint64_t* n_emitted) {
    auto ps = h->solver->particleSystemData();
    size_t before = ps->numberOfParticles();
    auto box = std::make_shared<Box3>(Vector3D(lower[0], lower[1], lower[2]), Vector3D(upper[0], upper[1], upper[2]));
    auto emitter = VolumeParticleEmitter3::builder()
                       .withSurface(box)
                       .withMaxRegion(h->solver->gridSystemData()->boundingBox())
                       .withSpacing(spacing)
                       .withInitialVelocity(Vector3D(v0[0], v0[1], v0[2]))
                       .withIsOneShot(true)
                       .withJitter(0.0)
                       .makeShared();
    emitter->setPointGenerator(std::make_shared<GridPointGenerator3>());  // as hybrid_liquid_sim does (main.cpp:250)
    emitter->setTarget(ps);
    emitter->update(0.0, 0.0);
    if (n_emitted) *n_emitted = ps->numberOfParticles() - before;
    return 0;
}

uint64_t jeto_hybrid_num_particles(const jeto_hybrid* h) { return h->solver->particleSystemData()->numberOfParticles(); }

int jeto_hybrid_get_particles(jeto_hybrid* h, double* pos, double* vel) {
    auto ps = h->solver->particleSystemData();
    size_t n = ps->numberOfParticles();
    auto P = ps->positions(); auto V = ps->velocities();
    for (size_t i = 0; i < n; ++i) {
        if (pos) { pos[3 * i] = P[i].x; pos[3 * i + 1] = P[i].y; pos[3 * i + 2] = P[i].z; }
        if (vel) { vel[3 * i] = V[i].x; vel[3 * i + 1] = V[i].y; vel[3 * i + 2] = V[i].z; }
    }
    return 0;
}

int jeto_hybrid_get_affine(jeto_hybrid* h, double* cx, double* cy, double* cz) {
    auto apic = std::dynamic_pointer_cast<ApicSolver3>(h->solver);
    if (!apic) return 1;
    size_t n = apic->_cX.size();
    for (size_t i = 0; i < n; ++i) {
        if (cx) { cx[3 * i] = apic->_cX[i].x; cx[3 * i + 1] = apic->_cX[i].y; cx[3 * i + 2] = apic->_cX[i].z; }
        if (cy) { cy[3 * i] = apic->_cY[i].x; cy[3 * i + 1] = apic->_cY[i].y; cy[3 * i + 2] = apic->_cY[i].z; }
        if (cz) { cz[3 * i] = apic->_cZ[i].x; cz[3 * i + 1] = apic->_cZ[i].y; cz[3 * i + 2] = apic->_cZ[i].z; }
    }
    return 0;
}

static size_t n3(const Size3& s) { return s.x * s.y * s.z; }

int jeto_hybrid_set_velocity(jeto_hybrid* h, const double* u, const double* v, const double* w) {
    auto vel = h->solver->gridSystemData()->velocity();
    if (u) std::memcpy(vel->uAccessor().data(), u, 8 * n3(vel->uSize()));
    if (v) std::memcpy(vel->vAccessor().data(), v, 8 * n3(vel->vSize()));
    if (w) std::memcpy(vel->wAccessor().data(), w, 8 * n3(vel->wSize()));
    return 0;
}

int jeto_hybrid_get_velocity(jeto_hybrid* h, double* u, double* v, double* w) {
    auto vel = h->solver->gridSystemData()->velocity();
    if (u) std::memcpy(u, vel->uAccessor().data(), 8 * n3(vel->uSize()));
    if (v) std::memcpy(v, vel->vAccessor().data(), 8 * n3(vel->vSize()));
    if (w) std::memcpy(w, vel->wAccessor().data(), 8 * n3(vel->wSize()));
    return 0;
}

int jeto_hybrid_get_sdf(jeto_hybrid* h, double* sdf) {
    auto g = h->solver->signedDistanceField();
    std::memcpy(sdf, g->dataAccessor().data(), 8 * n3(g->dataSize()));
    return 0;
}

int jeto_hybrid_get_pressure(jeto_hybrid* h, double* p) {
    const auto& x = h->pressure->pressure();
    if (n3(x.size()) == 0) return 1;
    std::memcpy(p, x.data(), 8 * n3(x.size()));
    return 0;
}

int jeto_hybrid_get_markers(jeto_hybrid* h, char* um, char* vm, char* wm) {
    auto& s = *h->solver;
    if (um) std::memcpy(um, s._uMarkers.data(), n3(s._uMarkers.size()));
    if (vm) std::memcpy(vm, s._vMarkers.data(), n3(s._vMarkers.size()));
    if (wm) std::memcpy(wm, s._wMarkers.data(), n3(s._wMarkers.size()));
    return 0;
}

int jeto_hybrid_cfl(jeto_hybrid* h, double dt, double* cfl) { *cfl = h->solver->cfl(dt); return 0; }
int jeto_hybrid_num_sub_steps(jeto_hybrid* h, double dt, uint32_t* n) { *n = h->solver->numberOfSubTimeSteps(dt); return 0; }

static void fillStats(jeto_hybrid* h, jeto_step_stats* st, double ms) {
    if (!st) return;
    std::memset(st, 0, sizeof(*st));
    if (h->iccg) {
        st->pcg_iterations = h->iccg->lastNumberOfIterations();
        st->pcg_last_residual = h->iccg->lastResidual();
        if (h->solver->useCompressedLinearSystem()) {
            const auto& cs = h->pressure->_compSystem;
            if (cs.b.size() > 0) {
                VectorND r(cs.b.size());
                FdmCompressedBlas3::residual(cs.A, cs.x, cs.b, &r);
                st->pcg_true_residual_l2 = FdmCompressedBlas3::l2Norm(r);
                st->pcg_b_l2 = FdmCompressedBlas3::l2Norm(cs.b);
            }
        } else {
            const auto& sys = h->pressure->_system;
            st->pcg_true_residual_l2 = trueResidual(sys.A, sys.x, sys.b);
            st->pcg_b_l2 = FdmBlas3::l2Norm(sys.b);
        }
        st->ms_total = ms;
        return;
    }
    st->pcg_iterations = h->lin->lastNumberOfIterations();
    st->pcg_last_residual = h->lin->lastResidual();
    const auto& sys = h->pressure->_mgSystem;
    if (!sys.A.levels.empty()) {
        st->pcg_true_residual_l2 = trueResidual(sys.A.levels.front(), sys.x.levels.front(), sys.b.levels.front());
        st->pcg_b_l2 = FdmBlas3::l2Norm(sys.b.levels.front());
    }
    st->ms_total = ms;
}

int jeto_hybrid_step(jeto_hybrid* h, double dt, jeto_step_stats* st) {
    double t0 = nowMs();
    h->solver->onAdvanceTimeStep(dt);
    fillStats(h, st, nowMs() - t0);
    return 0;
}

int jeto_hybrid_advance_frame(jeto_hybrid* h, double frame_dt, uint32_t* n_sub, jeto_step_stats* st) {
    // Same loop as PhysicsAnimation::advanceTimeStep's adaptive branch, driven through the solver's own
    // numberOfSubTimeSteps/onAdvanceTimeStep.
    double remaining = frame_dt;
    uint32_t count = 0;
    double t0 = nowMs();
    while (remaining > kEpsilonD) {
        unsigned int numSteps = h->solver->numberOfSubTimeSteps(remaining);
        double actual = remaining / static_cast<double>(numSteps);
        h->solver->onAdvanceTimeStep(actual);
        remaining -= actual;
        ++count;
    }
    if (n_sub) *n_sub = count;
    fillStats(h, st, nowMs() - t0);
    return 0;
}

int jeto_hybrid_apply_boundary_condition(jeto_hybrid* h) { h->solver->applyBoundaryCondition(); return 0; }
int jeto_hybrid_transfer_particles_to_grids(jeto_hybrid* h) { h->solver->transferFromParticlesToGrids(); return 0; }
int jeto_hybrid_build_signed_distance_field(jeto_hybrid* h) { h->solver->buildSignedDistanceField(); return 0; }
int jeto_hybrid_extrapolate_velocity_to_air(jeto_hybrid* h) { h->solver->extrapolateVelocityToAir(); return 0; }
int jeto_hybrid_begin_advance_time_step(jeto_hybrid* h, double dt) { h->solver->beginAdvanceTimeStep(dt); return 0; }
int jeto_hybrid_compute_external_forces(jeto_hybrid* h, double dt) { h->solver->computeExternalForces(dt); return 0; }
int jeto_hybrid_compute_pressure(jeto_hybrid* h, double dt, jeto_mgpcg_result* r) {
    h->solver->computePressure(dt);
    if (r) {
        jeto_step_stats st; fillStats(h, &st, 0);
        r->iterations = st.pcg_iterations; r->last_residual = st.pcg_last_residual;
        r->true_residual_l2 = st.pcg_true_residual_l2; r->b_l2 = st.pcg_b_l2; r->returned = 1; r->solve_ms = 0;
    }
    return 0;
}
int jeto_hybrid_transfer_grids_to_particles(jeto_hybrid* h) { h->solver->transferFromGridsToParticles(); return 0; }
int jeto_hybrid_move_particles(jeto_hybrid* h, double dt) { h->solver->moveParticles(dt); return 0; }
int jeto_hybrid_compute_advection(jeto_hybrid* h, double dt) { h->solver->computeAdvection(dt); return 0; }

// ------------------------------------------------------------------ GridBackwardEulerDiffusionSolver3
int jeto_hybrid_set_viscosity(jeto_hybrid* h, double coefficient) { h->solver->setViscosityCoefficient(coefficient); return 0; }

static void fillIccgResult(const std::shared_ptr<FdmIccgSolver3>& iccg, jeto_mgpcg_result* r) {
    std::memset(r, 0, sizeof(*r));
    r->iterations = iccg->lastNumberOfIterations();
    r->last_residual = iccg->lastResidual();
    r->returned = (iccg->lastResidual() <= iccg->tolerance() || iccg->lastNumberOfIterations() < iccg->maxNumberOfIterations()) ? 1 : 0;
}

int jeto_hybrid_compute_viscosity(jeto_hybrid* h, double dt, jeto_mgpcg_result results[3]) {
    h->solver->computeViscosity(dt);
    if (results) {
        auto diff = std::dynamic_pointer_cast<GridBackwardEulerDiffusionSolver3>(h->solver->diffusionSolver());
        auto iccg = diff ? std::dynamic_pointer_cast<FdmIccgSolver3>(diff->_systemSolver) : nullptr;
        for (int q = 0; q < 3; ++q) {
            std::memset(&results[q], 0, sizeof(results[q]));
            if (iccg && h->solver->viscosityCoefficient() > kEpsilonD) fillIccgResult(iccg, &results[q]);
        }
    }
    return 0;
}

static ScalarField3Ptr cellField(const jeto_grid* g, const double* data, double constant) {
    if (!data) return std::make_shared<ConstantScalarField3>(constant);
    auto grid = std::make_shared<CellCenteredScalarGrid3>(Size3(g->nx, g->ny, g->nz), Vector3D(g->hx, g->hy, g->hz),
                                                          Vector3D(g->ox, g->oy, g->oz));
    std::memcpy(grid->dataAccessor().data(), data, sizeof(double) * g->nx * g->ny * g->nz);
    return grid;
}

int jeto_diffusion_solve_faces(const jeto_grid* g, const double* u, const double* v, const double* w, double coefficient, double dt,
                               int boundary_type, const double* collider_sdf, const double* fluid_sdf, uint32_t max_iterations,
                               double tolerance, double* ou, double* ov, double* ow, jeto_mgpcg_result results[3]) {
    const Size3 res(g->nx, g->ny, g->nz);
    const Vector3D hh(g->hx, g->hy, g->hz), o(g->ox, g->oy, g->oz);
    FaceCenteredGrid3 src(res, hh, o), dst(res, hh, o);
    std::memcpy(src.uAccessor().data(), u, sizeof(double) * (g->nx + 1) * g->ny * g->nz);
    std::memcpy(src.vAccessor().data(), v, sizeof(double) * g->nx * (g->ny + 1) * g->nz);
    std::memcpy(src.wAccessor().data(), w, sizeof(double) * g->nx * g->ny * (g->nz + 1));
    // the collider SDF the fluid solvers pass is always a grid (kMaxD-filled when there is no collider)
    std::vector<double> none;
    if (!collider_sdf) { none.assign((size_t)g->nx * g->ny * g->nz, kMaxD); collider_sdf = none.data(); }
    auto boundary = cellField(g, collider_sdf, kMaxD);
    auto fluid = cellField(g, fluid_sdf, -kMaxD);
    GridBackwardEulerDiffusionSolver3 solver(boundary_type == 0 ? GridBackwardEulerDiffusionSolver3::Dirichlet
                                                                : GridBackwardEulerDiffusionSolver3::Neumann);
    auto iccg = std::make_shared<FdmIccgSolver3>(max_iterations, tolerance);
    solver.setLinearSystemSolver(iccg);
    solver.solve(src, coefficient, dt, &dst, *boundary, *fluid);
    std::memcpy(ou, dst.uAccessor().data(), sizeof(double) * (g->nx + 1) * g->ny * g->nz);
    std::memcpy(ov, dst.vAccessor().data(), sizeof(double) * g->nx * (g->ny + 1) * g->nz);
    std::memcpy(ow, dst.wAccessor().data(), sizeof(double) * g->nx * g->ny * (g->nz + 1));
    if (results) for (int q = 0; q < 3; ++q) fillIccgResult(iccg, &results[q]);   // the w solve (the class keeps its last only)
    return 0;
}

int jeto_diffusion_solve_cells(const jeto_grid* g, const double* source, double coefficient, double dt, int boundary_type,
                               const double* collider_sdf, const double* fluid_sdf, uint32_t max_iterations, double tolerance,
                               double* dest, char* markers, jeto_mgpcg_result* result) {
    const Size3 res(g->nx, g->ny, g->nz);
    const Vector3D hh(g->hx, g->hy, g->hz), o(g->ox, g->oy, g->oz);
    CellCenteredScalarGrid3 src(res, hh, o), dst(res, hh, o);
    const size_t n = (size_t)g->nx * g->ny * g->nz;
    std::memcpy(src.dataAccessor().data(), source, sizeof(double) * n);
    std::vector<double> none;
    if (!collider_sdf) { none.assign(n, kMaxD); collider_sdf = none.data(); }
    auto boundary = cellField(g, collider_sdf, kMaxD);
    auto fluid = cellField(g, fluid_sdf, -kMaxD);
    GridBackwardEulerDiffusionSolver3 solver(boundary_type == 0 ? GridBackwardEulerDiffusionSolver3::Dirichlet
                                                                : GridBackwardEulerDiffusionSolver3::Neumann);
    auto iccg = std::make_shared<FdmIccgSolver3>(max_iterations, tolerance);
    solver.setLinearSystemSolver(iccg);
    solver.solve(src, coefficient, dt, &dst, *boundary, *fluid);
    std::memcpy(dest, dst.dataAccessor().data(), sizeof(double) * n);
    if (markers) for (size_t i = 0; i < n; ++i) markers[i] = solver._markers.data()[i];
    if (result) fillIccgResult(iccg, result);
    return 0;
}

// ------------------------------------------------------------------ smoke solver
struct jeto_smoke {
    std::shared_ptr<GridSmokeSolver3> solver;
    std::shared_ptr<FdmMgpcgSolver3> lin;
    std::shared_ptr<GridFractionalSinglePhasePressureSolver3> pressure;
    bool solved = false;   // pressure() dereferences an empty MG hierarchy before the first solve
};

int jeto_smoke_create(const jeto_smoke_params* p, jeto_smoke** out) {
    auto* h = new jeto_smoke();
    Size3 res(p->grid.nx, p->grid.ny, p->grid.nz);
    Vector3D hh(p->grid.hx, p->grid.hy, p->grid.hz), o(p->grid.ox, p->grid.oy, p->grid.oz);
    h->solver = std::make_shared<GridSmokeSolver3>(res, hh, o);
    h->lin = makeMgpcg(&p->mgpcg);
    h->pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
    h->pressure->setLinearSystemSolver(h->lin);
    h->solver->setPressureSolver(h->pressure);
    h->solver->setGravity(Vector3D(p->gravity[0], p->gravity[1], p->gravity[2]));
    h->solver->setMaxCfl(p->max_cfl);
    h->solver->setClosedDomainBoundaryFlag(p->closed_domain_boundary_flag);
    h->solver->setUseCompressedLinearSystem(false);
    h->solver->setBuoyancySmokeDensityFactor(p->buoyancy_density_factor);
    h->solver->setBuoyancyTemperatureFactor(p->buoyancy_temperature_factor);
    h->solver->setSmokeDecayFactor(p->density_decay);
    h->solver->setTemperatureDecayFactor(p->temperature_decay);
    if (p->advection == 1) h->solver->setAdvectionSolver(std::make_shared<SemiLagrangian3>());   // else the default: CubicSemiLagrangian3
    *out = h;
    return 0;
}

int jeto_smoke_destroy(jeto_smoke* h) { delete h; return 0; }

int jeto_smoke_set_fields(jeto_smoke* h, const double* den, const double* temp, const double* u, const double* v, const double* w) {
    auto grids = h->solver->gridSystemData();
    auto res = grids->resolution();
    size_t nc = res.x * res.y * res.z;
    if (den) std::memcpy(h->solver->smokeDensity()->dataAccessor().data(), den, nc * 8);
    if (temp) std::memcpy(h->solver->temperature()->dataAccessor().data(), temp, nc * 8);
    auto vel = grids->velocity();
    if (u) std::memcpy(vel->uAccessor().data(), u, (res.x + 1) * res.y * res.z * 8);
    if (v) std::memcpy(vel->vAccessor().data(), v, res.x * (res.y + 1) * res.z * 8);
    if (w) std::memcpy(vel->wAccessor().data(), w, res.x * res.y * (res.z + 1) * 8);
    return 0;
}

int jeto_smoke_get_fields(jeto_smoke* h, double* den, double* temp, double* u, double* v, double* w, double* pr) {
    auto grids = h->solver->gridSystemData();
    auto res = grids->resolution();
    size_t nc = res.x * res.y * res.z;
    if (den) std::memcpy(den, h->solver->smokeDensity()->constDataAccessor().data(), nc * 8);
    if (temp) std::memcpy(temp, h->solver->temperature()->constDataAccessor().data(), nc * 8);
    auto vel = grids->velocity();
    if (u) std::memcpy(u, vel->uConstAccessor().data(), (res.x + 1) * res.y * res.z * 8);
    if (v) std::memcpy(v, vel->vConstAccessor().data(), res.x * (res.y + 1) * res.z * 8);
    if (w) std::memcpy(w, vel->wConstAccessor().data(), res.x * res.y * (res.z + 1) * 8);
    if (pr) {
        std::memset(pr, 0, nc * 8);
        if (h->solved) {
            const auto& x = h->pressure->pressure();
            if (x.size().x * x.size().y * x.size().z == nc) std::memcpy(pr, x.data(), nc * 8);
        }
    }
    return 0;
}

static void smokeResult(jeto_smoke* h, jeto_mgpcg_result* r) {
    if (!r) return;
    std::memset(r, 0, sizeof(*r));
    r->iterations = h->lin->lastNumberOfIterations();
    r->last_residual = h->lin->lastResidual();
    r->returned = 1;
    const auto& sys = h->pressure->_mgSystem;
    if (!sys.A.levels.empty()) {
        r->true_residual_l2 = trueResidual(sys.A.levels.front(), sys.x.levels.front(), sys.b.levels.front());
        r->b_l2 = FdmBlas3::l2Norm(sys.b.levels.front());
    }
}

int jeto_smoke_step(jeto_smoke* h, double dt, jeto_mgpcg_result* r) {
    h->solver->onAdvanceTimeStep(dt);
    h->solved = true;
    smokeResult(h, r);
    return 0;
}
int jeto_smoke_compute_external_forces(jeto_smoke* h, double dt) { h->solver->computeExternalForces(dt); return 0; }
int jeto_smoke_compute_pressure(jeto_smoke* h, double dt, jeto_mgpcg_result* r) {
    h->solver->computePressure(dt);
    h->solved = true;
    smokeResult(h, r);
    return 0;
}
int jeto_smoke_compute_advection(jeto_smoke* h, double dt) { h->solver->computeAdvection(dt); return 0; }
int jeto_smoke_end_advance_time_step(jeto_smoke* h, double dt) { h->solver->endAdvanceTimeStep(dt); return 0; }
int jeto_smoke_set_diffusion(jeto_smoke* h, double viscosity, double smokeDiffusion, double temperatureDiffusion) {
    h->solver->setViscosityCoefficient(viscosity);
    h->solver->setSmokeDiffusionCoefficient(smokeDiffusion);
    h->solver->setTemperatureDiffusionCoefficient(temperatureDiffusion);
    return 0;
}
int jeto_smoke_compute_viscosity(jeto_smoke* h, double dt) { h->solver->computeViscosity(dt); return 0; }

int jeto_smoke_emit_step(jeto_smoke* h, const jeto_collider_shape* shapes, uint32_t n, int as_set, double dmin, double dmax,
                         double tmin, double tmax) {
    if (n == 0 || (n != 1 && !as_set)) return 1;
    Surface3Ptr surface;
    if (as_set) {
        std::vector<Surface3Ptr> members;
        for (uint32_t i = 0; i < n; ++i) {
            auto m = makeShape(shapes[i]);
            if (!m) return 1;
            members.push_back(m);
        }
        surface = ImplicitSurfaceSet3::builder().withExplicitSurfaces(members).makeShared();
    } else {
        surface = makeShape(shapes[0]);
        if (!surface) return 1;
    }
    auto emitter = VolumeGridEmitter3::builder().withSourceRegion(surface).withIsOneShot(true).makeShared();   // smoke_sim/main.cpp:209-216
    emitter->addStepFunctionTarget(h->solver->smokeDensity(), dmin, dmax);
    emitter->addStepFunctionTarget(h->solver->temperature(), tmin, tmax);
    emitter->update(0.0, 0.0);
    return 0;
}

int jeto_smoke_set_collider_shapes(jeto_smoke* h, const jeto_collider_shape* shapes, uint32_t n, int as_set, const double lin[3],
                                   const double ang[3], double friction) {
    if (n == 0) { h->solver->setCollider(nullptr); return 0; }
    if (n != 1 && !as_set) return 1;
    Surface3Ptr surface;
    if (as_set) {
        std::vector<Surface3Ptr> members;
        for (uint32_t i = 0; i < n; ++i) {
            auto m = makeShape(shapes[i]);
            if (!m) return 1;
            members.push_back(m);
        }
        surface = std::make_shared<ImplicitSurfaceSet3>(members);
    } else {
        surface = makeShape(shapes[0]);
        if (!surface) return 1;
    }
    auto col = std::make_shared<RigidBodyCollider3>(surface, lin ? Vector3D(lin[0], lin[1], lin[2]) : Vector3D(),
                                                    ang ? Vector3D(ang[0], ang[1], ang[2]) : Vector3D());
    col->setFrictionCoefficient(friction);
    h->solver->setCollider(col);
    // what GridFluidSolver3::beginAdvanceTimeStep does before anything reads the collider grid (:448-453)
    auto grids = h->solver->gridSystemData();
    h->solver->_boundaryConditionSolver->updateCollider(col, grids->resolution(), grids->gridSpacing(), grids->origin());
    return 0;
}

int jeto_smoke_cfl(jeto_smoke* h, double dt, double* cfl) { *cfl = h->solver->cfl(dt); return 0; }
int jeto_smoke_num_sub_steps(jeto_smoke* h, double dt, uint32_t* n) { *n = h->solver->numberOfSubTimeSteps(dt); return 0; }

}  // extern "C"
