// jetcu_jet_adapters.h — header-only C++11 adapters that put the jetcu C ABI (include/jetcu.h) behind Jet's own
// virtual interfaces, so UNMODIFIED Jet host code (doyubkim/fluid-engine-dev) runs the hybrid-liquid hot path on a
// B200.  Compile this header inside a Jet application: it includes <jet/jet.h> from the Jet tree and links
// libjetcu.so.  Nothing of Jet is copied here; the adapters only subclass / call its public API.
//
//   seam 1  jetcu::CudaFdmMgpcgSolver3 : jet::FdmMgSolver3            (include/jet/fdm_mg_solver3.h:45)
//           install with  pressureSolver->setLinearSystemSolver(std::make_shared<jetcu::CudaFdmMgpcgSolver3>(...))
//           — Jet builds the multigrid system on the host, the GPU solves it (jetcu_mgpcg_solve).
//   seam 2  jetcu::CudaGridFractionalSinglePhasePressureSolver3 : jet::GridPressureSolver3
//           (include/jet/grid_pressure_solver3.h:58-74); install with  solver->setPressureSolver(...)
//           — weights, system build, MGPCG and the pressure gradient all run on the GPU (jetcu_pressure_solve).
//   seam 3  jetcu::CudaPicSolver3 / CudaFlipSolver3 / CudaApicSolver3 : jet::PicSolver3 / FlipSolver3 / ApicSolver3
//           — the whole sub-step (GridFluidSolver3::onAdvanceTimeStep, src/jet/grid_fluid_solver3.cpp:179-210)
//           runs device-resident (jetcu_hybrid_step); particles and grids are mirrored back into Jet's
//           ParticleSystemData3 / FaceCenteredGrid3 / signedDistanceField() objects.
//
// Error convention: the C ABI returns int codes; the adapters rethrow JETCU_ERR_INVALID_ARG as
// std::invalid_argument (what JET_THROW_INVALID_ARG_IF raises, include/jet/macros.h:37-41) and everything else as
// std::runtime_error.  There is no CPU fallback: without a usable B200 every call throws.
#ifndef JETCU_JET_ADAPTERS_H_
#define JETCU_JET_ADAPTERS_H_

#include <jet/jet.h>

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "../include/jetcu.h"

namespace jetcu {

// ------------------------------------------------------------------------------------------------ context
class Context {
 public:
    explicit Context(int device = 0) {
        int rc = jetcu_ctx_create(device, &_ctx);
        if (rc != JETCU_OK) {
            const char* msg = jetcu_last_error(nullptr);
            throw std::runtime_error(std::string("jetcu_ctx_create failed: ") + (msg ? msg : "?"));
        }
    }
    ~Context() { jetcu_ctx_destroy(_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    jetcu_ctx* get() const { return _ctx; }

    void check(int rc) const {
        if (rc == JETCU_OK) return;
        const char* msg = jetcu_last_error(_ctx);
        std::string text = std::string("jetcu: ") + (msg ? msg : "error");
        if (rc == JETCU_ERR_INVALID_ARG) throw std::invalid_argument(text);
        throw std::runtime_error(text);
    }

    //! One lazily created context per process for device 0 (Jet's solvers are single-threaded per object).
    static std::shared_ptr<Context> shared(int device = 0) {
        static std::shared_ptr<Context> inst[16];
        if (device < 0 || device >= 16) throw std::invalid_argument("jetcu: device index out of range");
        if (!inst[device]) inst[device] = std::make_shared<Context>(device);
        return inst[device];
    }

 private:
    jetcu_ctx* _ctx = nullptr;
};
typedef std::shared_ptr<Context> ContextPtr;

inline jetcu_mgpcg_params makeMgpcgParams(unsigned int, size_t, unsigned int, unsigned int, unsigned int, unsigned int, double, double, bool);
//! == constructor arguments of jet::FdmIccgSolver3 (Jet's default linear solver: 100 iterations, 1e-6)
inline jetcu_mgpcg_params makeIccgParams(unsigned int maxNumberOfIterations = 100, double tolerance = 1e-6) {
    jetcu_mgpcg_params p = makeMgpcgParams(maxNumberOfIterations, 1, 0, 0, 0, 0, tolerance, 1.0, true);
    p.linear_solver = JETCU_SOLVER_ICCG;
    return p;
}

inline jetcu_mgpcg_params makeMgpcgParams(unsigned int numberOfCgIter, size_t maxNumberOfLevels,
                                          unsigned int numberOfRestrictionIter, unsigned int numberOfCorrectionIter,
                                          unsigned int numberOfCoarsestIter, unsigned int numberOfFinalIter,
                                          double maxTolerance, double sorFactor, bool useRedBlackOrdering) {
    jetcu_mgpcg_params p;
    std::memset(&p, 0, sizeof(p));
    p.max_iterations = numberOfCgIter;
    p.max_levels = static_cast<uint32_t>(maxNumberOfLevels);
    p.n_restriction = numberOfRestrictionIter;
    p.n_correction = numberOfCorrectionIter;
    p.n_coarsest = numberOfCoarsestIter;
    p.n_final = numberOfFinalIter;
    p.tolerance = maxTolerance;
    p.sor_factor = sorFactor;
    p.use_red_black = useRedBlackOrdering ? 1 : 0;
    return p;
}

// ------------------------------------------------------------------------------------------------ seam 1
//! Drop-in for jet::FdmMgpcgSolver3 (which is `final`, include/jet/fdm_mgpcg_solver3.h:23): same constructor
//! arguments, same accessors; GridFractionalSinglePhasePressureSolver3 routes to solve(FdmMgLinearSystem3*) for any
//! FdmMgSolver3 subclass (src/jet/grid_fractional_single_phase_pressure_solver3.cpp:477-478,497-510).
class CudaFdmMgpcgSolver3 final : public jet::FdmMgSolver3 {
 public:
    CudaFdmMgpcgSolver3(unsigned int numberOfCgIter, size_t maxNumberOfLevels,
                        unsigned int numberOfRestrictionIter = 5, unsigned int numberOfCorrectionIter = 5,
                        unsigned int numberOfCoarsestIter = 20, unsigned int numberOfFinalIter = 20,
                        double maxTolerance = 1e-9, double sorFactor = 1.5, bool useRedBlackOrdering = false,
                        ContextPtr ctx = ContextPtr())
        : jet::FdmMgSolver3(maxNumberOfLevels, numberOfRestrictionIter, numberOfCorrectionIter, numberOfCoarsestIter,
                            numberOfFinalIter, maxTolerance, sorFactor, useRedBlackOrdering),
          _ctx(ctx ? ctx : Context::shared()),
          _p(makeMgpcgParams(numberOfCgIter, maxNumberOfLevels, numberOfRestrictionIter, numberOfCorrectionIter,
                             numberOfCoarsestIter, numberOfFinalIter, maxTolerance, sorFactor, useRedBlackOrdering)) {
        std::memset(&_last, 0, sizeof(_last));
    }

    bool solve(jet::FdmMgLinearSystem3* system) override {
        const size_t levels = system->numberOfLevels();
        if (levels == 0) return true;
        static_assert(sizeof(jet::FdmMatrixRow3) == sizeof(jetcu_matrix_row), "FdmMatrixRow3 must be 4 doubles");
        std::vector<const jetcu_matrix_row*> A(levels);
        for (size_t l = 0; l < levels; ++l)
            A[l] = reinterpret_cast<const jetcu_matrix_row*>(system->A.levels[l].data());
        const jet::Size3 n = system->A.levels[0].size();
        JET_THROW_INVALID_ARG_IF(system->b.levels[0].size() != n);
        JET_THROW_INVALID_ARG_IF(system->x.levels[0].size() != n);
        _ctx->check(jetcu_mgpcg_solve(_ctx->get(), n.x, n.y, n.z, static_cast<uint32_t>(levels), A.data(),
                                      system->b.levels[0].data(), system->x.levels[0].data(), &_p, &_last));
        return _last.returned != 0;
    }

    unsigned int maxNumberOfIterations() const { return _p.max_iterations; }
    unsigned int lastNumberOfIterations() const { return _last.iterations; }
    double tolerance() const { return _p.tolerance; }
    double lastResidual() const { return _last.last_residual; }
    //! ||b - A x||_2 recomputed at exit and ||b||_2 (not available in Jet).
    double lastTrueResidual() const { return _last.true_residual_l2; }
    double lastRhsNorm() const { return _last.b_l2; }
    //! Device time of the last solve in milliseconds (CUDA events).
    double lastSolveMilliseconds() const { return _last.solve_ms; }
    const jetcu_mgpcg_params& cudaParams() const { return _p; }

 private:
    ContextPtr _ctx;
    jetcu_mgpcg_params _p;
    jetcu_mgpcg_result _last;
};
typedef std::shared_ptr<CudaFdmMgpcgSolver3> CudaFdmMgpcgSolver3Ptr;

//! Drop-in for jet::FdmIccgSolver3 (`final`, include/jet/fdm_iccg_solver3.h:18), the default linear solver of Jet's pressure and
//! diffusion solvers: same constructor arguments, same accessors.  solve(FdmLinearSystem3*) runs on the device
//! (jetcu_iccg_solve: incomplete-Cholesky substitutions as tile-hyperplane sweeps).  solveCompressed is NOT taken over — the
//! base class's `return false` stays, which makes GridFractionalSinglePhasePressureSolver3 report failure rather than use
//! another code path; install it with setUseCompressedLinearSystem(false), or use the seam-2 / seam-3 classes, which accept the
//! compressed flag and solve the equivalent uncompressed system on the device.
class CudaFdmIccgSolver3 final : public jet::FdmLinearSystemSolver3 {
 public:
    CudaFdmIccgSolver3(unsigned int maxNumberOfIterations, double tolerance, ContextPtr ctx = ContextPtr())
        : _ctx(ctx ? ctx : Context::shared()), _maxIter(maxNumberOfIterations), _tol(tolerance) {
        std::memset(&_last, 0, sizeof(_last));
    }

    bool solve(jet::FdmLinearSystem3* system) override {
        const jet::Size3 n = system->A.size();
        JET_THROW_INVALID_ARG_IF(system->b.size() != n);
        JET_THROW_INVALID_ARG_IF(system->x.size() != n);
        if (n.x * n.y * n.z == 0) return true;
        static_assert(sizeof(jet::FdmMatrixRow3) == sizeof(jetcu_matrix_row), "FdmMatrixRow3 must be 4 doubles");
        _ctx->check(jetcu_iccg_solve(_ctx->get(), n.x, n.y, n.z, reinterpret_cast<const jetcu_matrix_row*>(system->A.data()),
                                     system->b.data(), system->x.data(), _maxIter, _tol, &_last));
        return _last.returned != 0;
    }

    unsigned int maxNumberOfIterations() const { return _maxIter; }
    unsigned int lastNumberOfIterations() const { return _last.iterations; }
    double tolerance() const { return _tol; }
    double lastResidual() const { return _last.last_residual; }
    double lastSolveMilliseconds() const { return _last.solve_ms; }

 private:
    ContextPtr _ctx;
    unsigned int _maxIter;
    double _tol;
    jetcu_mgpcg_result _last;
};
typedef std::shared_ptr<CudaFdmIccgSolver3> CudaFdmIccgSolver3Ptr;

namespace detail {

//! MGPCG parameters of whatever linear solver a Jet pressure solver carries (FdmMgpcgSolver3 or our seam-1 class).
inline bool mgpcgParamsOf(const jet::FdmLinearSystemSolver3Ptr& lin, jetcu_mgpcg_params* out) {
    if (auto s = std::dynamic_pointer_cast<jet::FdmMgpcgSolver3>(lin)) {
        const auto& mp = s->params();
        *out = makeMgpcgParams(s->maxNumberOfIterations(), mp.maxNumberOfLevels, mp.numberOfRestrictionIter,
                               mp.numberOfCorrectionIter, mp.numberOfCoarsestIter, mp.numberOfFinalIter,
                               s->tolerance(), s->sorFactor(), s->useRedBlackOrdering());
        return true;
    }
    if (auto s = std::dynamic_pointer_cast<CudaFdmMgpcgSolver3>(lin)) {
        *out = s->cudaParams();
        return true;
    }
    return false;
}

//! Parameters of an ICCG linear solver (Jet's FdmIccgSolver3 or our CudaFdmIccgSolver3) as jetcu parameters.
inline bool iccgParamsOf(const jet::FdmLinearSystemSolver3Ptr& lin, jetcu_mgpcg_params* out) {
    unsigned int it = 0;
    double tol = 0.0;
    if (auto s = std::dynamic_pointer_cast<jet::FdmIccgSolver3>(lin)) { it = s->maxNumberOfIterations(); tol = s->tolerance(); }
    else if (auto c = std::dynamic_pointer_cast<CudaFdmIccgSolver3>(lin)) { it = c->maxNumberOfIterations(); tol = c->tolerance(); }
    else return false;
    *out = makeMgpcgParams(it, 1, 0, 0, 0, 0, tol, 1.0, true);
    out->linear_solver = JETCU_SOLVER_ICCG;
    return true;
}

//! Cell-centred samples of a ScalarField3 for grid (res, h, origin): the raw array when the field already is a
//! matching CellCenteredScalarGrid3 (what Jet's solvers pass: SURVEY §8b), nullptr for the +-kMaxD constants the
//! interface defaults to, otherwise host sampling at the cell centres into `scratch`.
inline const double* cellSamples(const jet::ScalarField3& f, const jet::Size3& res, const jet::Vector3D& h,
                                 const jet::Vector3D& origin, bool nullWhenPositiveConstant,
                                 std::vector<double>* scratch) {
    if (auto g = dynamic_cast<const jet::CellCenteredScalarGrid3*>(&f)) {
        if (g->resolution() == res && g->gridSpacing().isSimilar(h, 0.0) && g->origin().isSimilar(origin, 0.0))
            return g->constDataAccessor().data();
    }
    if (auto c = dynamic_cast<const jet::ConstantScalarField3*>(&f)) {
        const double v = c->sample(jet::Vector3D());
        if (nullWhenPositiveConstant ? (v >= jet::kMaxD) : (v <= -jet::kMaxD)) return nullptr;
    }
    scratch->resize(res.x * res.y * res.z);
    double* out = scratch->data();
    jet::parallelFor(jet::kZeroSize, res.x, jet::kZeroSize, res.y, jet::kZeroSize, res.z,
                     [&](size_t i, size_t j, size_t k) {
                         jet::Vector3D x = origin + jet::Vector3D((i + 0.5) * h.x, (j + 0.5) * h.y, (k + 0.5) * h.z);
                         out[(k * res.y + j) * res.x + i] = f.sample(x);
                     });
    return out;
}

// ---- analytic surfaces (colliders, emitters): Jet objects -> jetcu_collider_shape records
inline bool identityTransform(const jet::Transform3& t) {
    const jet::QuaternionD q = t.orientation();
    return t.translation().lengthSquared() == 0.0 && q.w == 1.0 && q.x == 0.0 && q.y == 0.0 && q.z == 0.0;
}

inline jetcu_collider_shape shapeOf(const jet::Surface3Ptr& surface) {
    if (!identityTransform(surface->transform))
        throw std::runtime_error("jetcu: device colliders need identity transforms");
    jetcu_collider_shape s;
    std::memset(&s, 0, sizeof(s));
    s.flipped = surface->isNormalFlipped ? 1 : 0;
    if (auto box = std::dynamic_pointer_cast<jet::Box3>(surface)) {
        s.kind = JETCU_COLLIDER_BOX;
        s.a[0] = box->bound.lowerCorner.x; s.a[1] = box->bound.lowerCorner.y; s.a[2] = box->bound.lowerCorner.z;
        s.b[0] = box->bound.upperCorner.x; s.b[1] = box->bound.upperCorner.y; s.b[2] = box->bound.upperCorner.z;
    } else if (auto sphere = std::dynamic_pointer_cast<jet::Sphere3>(surface)) {
        s.kind = JETCU_COLLIDER_SPHERE;
        s.a[0] = sphere->center.x; s.a[1] = sphere->center.y; s.a[2] = sphere->center.z;
        s.b[0] = sphere->radius;
    } else if (auto cyl = std::dynamic_pointer_cast<jet::Cylinder3>(surface)) {
        s.kind = JETCU_COLLIDER_CYLINDER;
        s.a[0] = cyl->center.x; s.a[1] = cyl->center.y; s.a[2] = cyl->center.z;
        s.b[0] = cyl->radius; s.b[1] = cyl->height;
    } else if (auto plane = std::dynamic_pointer_cast<jet::Plane3>(surface)) {
        s.kind = JETCU_COLLIDER_PLANE;
        s.a[0] = plane->point.x; s.a[1] = plane->point.y; s.a[2] = plane->point.z;
        s.b[0] = plane->normal.x; s.b[1] = plane->normal.y; s.b[2] = plane->normal.z;
    } else {
        throw std::runtime_error("jetcu: device colliders are Box3 / Sphere3 / Cylinder3 / Plane3 (closed forms) or an "
                                 "ImplicitSurfaceSet3 of them; got another Surface3");
    }
    return s;
}

//! One explicit surface, or the members of an ImplicitSurfaceSet3 (asSet = 1), as jetcu_collider_shape records.
inline void collectShapes(const jet::Surface3Ptr& surface, std::vector<jetcu_collider_shape>* shapes, int* asSet) {
    shapes->clear();
    *asSet = 0;
    if (auto set = std::dynamic_pointer_cast<jet::ImplicitSurfaceSet3>(surface)) {
        if (set->isNormalFlipped || !identityTransform(set->transform))
            throw std::runtime_error("jetcu: a device surface set needs an identity transform and unflipped normals");
        if (set->numberOfSurfaces() == 0 || set->numberOfSurfaces() > JETCU_COLLIDER_MAX_SHAPES)
            throw std::runtime_error("jetcu: a device surface set holds 1.." + std::to_string(JETCU_COLLIDER_MAX_SHAPES) + " surfaces");
        for (size_t i = 0; i < set->numberOfSurfaces(); ++i) {
            auto wrap = std::dynamic_pointer_cast<jet::SurfaceToImplicit3>(set->surfaceAt(i));
            if (!wrap || wrap->isNormalFlipped || !identityTransform(wrap->transform))
                throw std::runtime_error("jetcu: surface set members must be explicit surfaces (addExplicitSurface)");
            shapes->push_back(shapeOf(wrap->surface()));
        }
        *asSet = 1;
    } else {
        shapes->push_back(shapeOf(surface));
    }
}


}  // namespace detail

// ------------------------------------------------------------------------------------------------ seam 2
//! Drop-in for jet::GridFractionalSinglePhasePressureSolver3 with an MGPCG linear solver.
class CudaGridFractionalSinglePhasePressureSolver3 final : public jet::GridPressureSolver3 {
 public:
    explicit CudaGridFractionalSinglePhasePressureSolver3(
        const jetcu_mgpcg_params& params = makeMgpcgParams(100, 6, 5, 5, 20, 20, 1e-9, 1.5, true),
        ContextPtr ctx = ContextPtr())
        : _ctx(ctx ? ctx : Context::shared()), _p(params) {
        std::memset(&_last, 0, sizeof(_last));
    }

    void solve(const jet::FaceCenteredGrid3& input, double timeIntervalInSeconds, jet::FaceCenteredGrid3* output,
               const jet::ScalarField3& boundarySdf = jet::ConstantScalarField3(jet::kMaxD),
               const jet::VectorField3& boundaryVelocity = jet::ConstantVectorField3({0, 0, 0}),
               const jet::ScalarField3& fluidSdf = jet::ConstantScalarField3(-jet::kMaxD),
               bool useCompressed = false) override {
        (void)timeIntervalInSeconds;
        // With a multigrid linear-system solver installed — and this path IS FdmMgpcgSolver3 — the reference neither builds nor
        // solves the compressed system: buildSystem fills _mgSystem whatever the flag says and solve calls
        // _mgSystemSolver->solve(&_mgSystem) (src/jet/grid_fractional_single_phase_pressure_solver3.cpp:465-481, 651-690).
        // So setUseCompressedLinearSystem(true), as Jet's examples do, changes nothing here either (checked on the CPU by
        // `dropin_check_next --what compressed`: stock Jet with MGPCG, flag on vs off, bit-identical particles).
        // With ICCG parameters (params.linear_solver = JETCU_SOLVER_ICCG, e.g. from makeIccgParams) the flag is forwarded: the
        // device then solves the reference's compressed system in its equivalent uncompressed form.
        jetcu_mgpcg_params prm = _p;
        if ((prm.linear_solver & JETCU_SOLVER_ICCG) && useCompressed) prm.linear_solver |= JETCU_SOLVER_COMPRESSED;
        const jet::Size3 res = input.resolution();
        JET_THROW_INVALID_ARG_IF(output->resolution() != res);
        const jet::Vector3D h = input.gridSpacing(), o = input.origin();
        jetcu_grid g;
        g.nx = res.x; g.ny = res.y; g.nz = res.z;
        g.hx = h.x; g.hy = h.y; g.hz = h.z;
        g.ox = o.x; g.oy = o.y; g.oz = o.z;

        const double* csdf = detail::cellSamples(boundarySdf, res, h, o, true, &_colliderScratch);
        const double* fsdf = detail::cellSamples(fluidSdf, res, h, o, false, &_fluidScratch);

        // boundaryVelocity is sampled at face positions, one component per face family
        // (src/jet/grid_fractional_single_phase_pressure_solver3.cpp:242-255); only needed next to a collider.
        const double *bu = nullptr, *bv = nullptr, *bw = nullptr;
        bool zeroVel = false;
        if (auto c = dynamic_cast<const jet::ConstantVectorField3*>(&boundaryVelocity))
            zeroVel = c->sample(jet::Vector3D()).lengthSquared() == 0.0;
        if (csdf != nullptr && !zeroVel) {
            _bu.resize((res.x + 1) * res.y * res.z);
            _bv.resize(res.x * (res.y + 1) * res.z);
            _bw.resize(res.x * res.y * (res.z + 1));
            auto uPos = input.uPosition();
            auto vPos = input.vPosition();
            auto wPos = input.wPosition();
            double *pu = _bu.data(), *pv = _bv.data(), *pw = _bw.data();
            jet::parallelFor(jet::kZeroSize, res.x + 1, jet::kZeroSize, res.y, jet::kZeroSize, res.z,
                             [&](size_t i, size_t j, size_t k) {
                                 pu[(k * res.y + j) * (res.x + 1) + i] = boundaryVelocity.sample(uPos(i, j, k)).x;
                             });
            jet::parallelFor(jet::kZeroSize, res.x, jet::kZeroSize, res.y + 1, jet::kZeroSize, res.z,
                             [&](size_t i, size_t j, size_t k) {
                                 pv[(k * (res.y + 1) + j) * res.x + i] = boundaryVelocity.sample(vPos(i, j, k)).y;
                             });
            jet::parallelFor(jet::kZeroSize, res.x, jet::kZeroSize, res.y, jet::kZeroSize, res.z + 1,
                             [&](size_t i, size_t j, size_t k) {
                                 pw[(k * res.y + j) * res.x + i] = boundaryVelocity.sample(wPos(i, j, k)).z;
                             });
            bu = pu; bv = pv; bw = pw;
        }

        _pressure.resize(res);
        _ctx->check(jetcu_pressure_solve(_ctx->get(), &g, input.uConstAccessor().data(), input.vConstAccessor().data(),
                                         input.wConstAccessor().data(), csdf, bu, bv, bw, fsdf, &prm,
                                         output->uAccessor().data(), output->vAccessor().data(),
                                         output->wAccessor().data(), _pressure.data(), &_last));
    }

    jet::GridBoundaryConditionSolver3Ptr suggestedBoundaryConditionSolver() const override {
        return std::make_shared<jet::GridFractionalBoundaryConditionSolver3>();
    }

    //! == GridFractionalSinglePhasePressureSolver3::pressure().
    const jet::FdmVector3& pressure() const { return _pressure; }
    const jetcu_mgpcg_result& lastResult() const { return _last; }
    const jetcu_mgpcg_params& cudaParams() const { return _p; }
    void setCudaParams(const jetcu_mgpcg_params& p) { _p = p; }

 private:
    ContextPtr _ctx;
    jetcu_mgpcg_params _p;
    jetcu_mgpcg_result _last;
    jet::FdmVector3 _pressure;
    std::vector<double> _colliderScratch, _fluidScratch, _bu, _bv, _bw;
};
typedef std::shared_ptr<CudaGridFractionalSinglePhasePressureSolver3> CudaGridFractionalSinglePhasePressureSolver3Ptr;

// ------------------------------------------------------------------------------------------------ seam 3
//! When the host-side Jet objects (ParticleSystemData3 positions/velocities, velocity grid, SDF grid) are refreshed
//! from the device.
enum class HostMirror {
    kEveryStep,  //!< after every sub-step: strict drop-in (any Jet code reading the solver's state sees it)
    kOnDemand    //!< only in syncToHost(): device-resident stepping without PCIe traffic per step
};

//! Device-resident PIC/FLIP/APIC solver behind Jet's class hierarchy.  Base = jet::PicSolver3 / FlipSolver3 /
//! ApicSolver3.  Overrides PhysicsAnimation::onAdvanceTimeStep / numberOfSubTimeSteps (include/jet/physics_animation.h:
//! 109,124), i.e. the whole of GridFluidSolver3::onAdvanceTimeStep; the particle emitter still runs on the host
//! (VolumeParticleEmitter3 etc.), newly emitted particles are appended to the device set.
template <class Base, int Mode>
class CudaHybridSolver3 : public Base {
 public:
    CudaHybridSolver3(const jet::Size3& resolution, const jet::Vector3D& gridSpacing, const jet::Vector3D& gridOrigin,
                      const jetcu_mgpcg_params& mgpcg = makeMgpcgParams(100, 6, 5, 5, 20, 20, 1e-9, 1.5, true),
                      ContextPtr ctx = ContextPtr())
        : Base(resolution, gridSpacing, gridOrigin), _ctx(ctx ? ctx : Context::shared()), _mgpcg(mgpcg) {
        // what runs on the device is what the Jet object shows: a fractional pressure solver carrying an FdmMgpcgSolver3 with
        // these parameters replaces the default-constructed one (whose linear solver is FdmIccgSolver3).  Callers may
        // install their own afterwards (setPressureSolver); ensureDevice() re-reads it every step.
        auto pressure = std::make_shared<jet::GridFractionalSinglePhasePressureSolver3>();
        pressure->setLinearSystemSolver(std::make_shared<jet::FdmMgpcgSolver3>(
            mgpcg.max_iterations, mgpcg.max_levels, mgpcg.n_restriction, mgpcg.n_correction, mgpcg.n_coarsest, mgpcg.n_final,
            mgpcg.tolerance, mgpcg.sor_factor, mgpcg.use_red_black != 0));
        this->setPressureSolver(pressure);
    }

    ~CudaHybridSolver3() override {
        if (_h) jetcu_hybrid_destroy(_h);
    }

    void setHostMirror(HostMirror m) { _mirror = m; }
    HostMirror hostMirror() const { return _mirror; }
    void setCudaMgpcgParams(const jetcu_mgpcg_params& p) { _mgpcg = p; }
    //! Opt in to running MGPCG (the constructor's parameters) although the Jet-side pressure solver carries another linear
    //! system solver (FdmIccgSolver3 by default).  Off by default: a different solver than the one configured is an error.
    void setAllowSolverSubstitution(bool allow) { _allowSolverSubstitution = allow; }
    bool allowSolverSubstitution() const { return _allowSolverSubstitution; }

    //! Tell the adapter that host-side particles or the velocity grid were edited by the caller since the last step.
    void markHostDirty() { _hostDirty = true; }

    //! Copies particles, velocity grid and SDF from the device into Jet's objects (kOnDemand mode).
    void syncToHost() {
        if (!_h) return;
        auto particles = this->particleSystemData();
        const size_t n = jetcu_hybrid_num_particles(_h);
        particles->resize(n);
        if (n > 0)
            _ctx->check(jetcu_hybrid_get_particles(_h, reinterpret_cast<double*>(particles->positions().data()),
                                                   reinterpret_cast<double*>(particles->velocities().data())));
        auto vel = this->velocity();
        _ctx->check(jetcu_hybrid_get_velocity(_h, vel->uAccessor().data(), vel->vAccessor().data(),
                                              vel->wAccessor().data()));
        auto sdf = this->signedDistanceField();
        _ctx->check(jetcu_hybrid_get_sdf(_h, sdf->dataAccessor().data()));
        _deviceCount = n;
    }

    const jetcu_step_stats& lastStepStats() const { return _stats; }

    //! saveParticleAsPos / saveParticleAsXyz of src/examples/hybrid_liquid_sim/main.cpp:29-63 (format JETCU_FRAME_POS /
    //! JETCU_FRAME_XYZ), byte-identical files, written from the device-resident particle set without refreshing the
    //! host mirror.
    void saveParticles(const std::string& path, int format) {
        ensureDevice();
        uploadIfNeeded();
        _ctx->check(jetcu_hybrid_save_positions(_h, path.c_str(), format));
    }

    //! Device-side equivalent of attaching a one-shot jet::VolumeParticleEmitter3 (GridPointGenerator3, jitter 0, maxRegion =
    //! the domain, as src/examples/hybrid_liquid_sim sets its emitters up) over `surface` and letting it fire: the lattice is
    //! generated, tested against the surface and appended on the GPU in the reference's emission order
    //! (jetcu_hybrid_emit_shapes) — no host lattice, no upload.  `surface`: Box3 / Sphere3 / Cylinder3 / Plane3 or an
    //! ImplicitSurfaceSet3 of them.  Returns the number of particles emitted.  kEveryStep refreshes the host mirror;
    //! kOnDemand only resizes it (values arrive with syncToHost()).
    size_t emitOnDevice(const jet::Surface3Ptr& surface, double spacing, const jet::Vector3D& initialVelocity = jet::Vector3D(),
                        int lattice = JETCU_LATTICE_GRID /* JETCU_LATTICE_BCC = VolumeParticleEmitter3's default generator */) {
        ensureDevice();
        uploadIfNeeded();
        std::vector<jetcu_collider_shape> shapes;
        int asSet = 0;
        detail::collectShapes(surface, &shapes, &asSet);
        const double v0[3] = {initialVelocity.x, initialVelocity.y, initialVelocity.z};
        uint64_t n = 0;
        _ctx->check(jetcu_hybrid_emit_shapes(_h, shapes.data(), (uint32_t)shapes.size(), asSet, lattice, spacing, v0, &n));
        _deviceCount += n;
        if (_mirror == HostMirror::kEveryStep) syncToHost();
        else this->particleSystemData()->resize(_deviceCount);
        return (size_t)n;
    }

 protected:
    void onAdvanceTimeStep(double timeIntervalInSeconds) override {
        auto grids = this->gridSystemData();
        if (grids->resolution().x == 0 || grids->resolution().y == 0 || grids->resolution().z == 0) return;
        // GridFluidSolver3::computeViscosity (grid_fluid_solver3.cpp:231-241) runs on the device with what a default-
        // constructed solver carries: GridBackwardEulerDiffusionSolver3 (Neumann) over FdmIccgSolver3(100, kEpsilonD).
        // Any other diffusion solver installed on the Jet object is refused, never replaced silently.
        if (this->viscosityCoefficient() > jet::kEpsilonD) {
            auto diff = std::dynamic_pointer_cast<jet::GridBackwardEulerDiffusionSolver3>(this->diffusionSolver());
            if (!diff)
                throw std::runtime_error("jetcu: computeViscosity on the device needs Jet's GridBackwardEulerDiffusionSolver3 "
                                         "(the solver's default); another GridDiffusionSolver3 is installed");
        }
        ensureDevice();
        _ctx->check(jetcu_hybrid_set_viscosity(_h, this->viscosityCoefficient() > jet::kEpsilonD ? this->viscosityCoefficient() : 0.0));
        // GridFluidSolver3::beginAdvanceTimeStep -> updateCollider (grid_fluid_solver3.cpp:255-275, 371-375): animated
        // colliders move in their onBeginUpdate callbacks, so it has to run before the collider is forwarded
        if (this->collider() != nullptr) this->collider()->update(this->currentTimeInSeconds(), timeIntervalInSeconds);
        syncCollider();
        // PicSolver3::onBeginAdvanceTimeStep -> updateParticleEmitter (src/jet/pic_solver3.cpp:63-75)
        if (this->particleEmitter() != nullptr)
            this->particleEmitter()->update(this->currentTimeInSeconds(), timeIntervalInSeconds);
        uploadIfNeeded();
        _ctx->check(jetcu_hybrid_step(_h, timeIntervalInSeconds, &_stats));
        if (_mirror == HostMirror::kEveryStep) syncToHost();
    }

    unsigned int numberOfSubTimeSteps(double timeIntervalInSeconds) const override {
        if (!_h || _hostDirty) return Base::numberOfSubTimeSteps(timeIntervalInSeconds);
        uint32_t n = 1;
        _ctx->check(jetcu_hybrid_num_sub_steps(_h, timeIntervalInSeconds, &n));
        return n;
    }

 private:
    void releaseDevice() {
        if (_h) jetcu_hybrid_destroy(_h);
        _h = nullptr;
        _deviceCount = 0;
        _hostDirty = true;
        _hasCollider = false;
    }

    //! Reads the solver's knobs from the Jet object EVERY step (setGravity / setMaxCfl / setClosedDomainBoundaryFlag /
    //! setPicBlendingFactor / a re-configured linear solver between steps are honoured like in GridFluidSolver3) and creates
    //! the device solver on first use.  The pressure solve on the device is MGPCG or ICCG, whichever the Jet-side pressure
    //! solver carries (FdmMgpcgSolver3 / FdmIccgSolver3 or their Cuda* counterparts); any other linear system solver is an
    //! error unless the caller opted in with setAllowSolverSubstitution(true) (then: the constructor's MGPCG parameters).
    void ensureDevice() {
        jetcu_hybrid_params p;
        jetcu_hybrid_default_params(&p);
        auto grids = this->gridSystemData();
        const jet::Size3 res = grids->resolution();
        const jet::Vector3D h = grids->gridSpacing(), o = grids->origin();
        p.grid.nx = res.x; p.grid.ny = res.y; p.grid.nz = res.z;
        p.grid.hx = h.x; p.grid.hy = h.y; p.grid.hz = h.z;
        p.grid.ox = o.x; p.grid.oy = o.y; p.grid.oz = o.z;
        p.mode = Mode;
        p.closed_domain_boundary_flag = this->closedDomainBoundaryFlag();
        p.pic_blending_factor = picBlend(static_cast<Base*>(this));
        const jet::Vector3D g = this->gravity();
        p.gravity[0] = g.x; p.gravity[1] = g.y; p.gravity[2] = g.z;
        p.max_cfl = this->maxCfl();
        jetcu_mgpcg_params mg = _mgpcg;
        auto cu = std::dynamic_pointer_cast<CudaGridFractionalSinglePhasePressureSolver3>(this->pressureSolver());
        auto frac = std::dynamic_pointer_cast<jet::GridFractionalSinglePhasePressureSolver3>(this->pressureSolver());
        jetcu_mgpcg_params fromJet;
        if (cu) mg = cu->cudaParams();
        else if (frac && detail::mgpcgParamsOf(frac->linearSystemSolver(), &fromJet)) mg = fromJet;
        else if (frac && detail::iccgParamsOf(frac->linearSystemSolver(), &fromJet)) {
            // Jet's default solver (what src/examples/hybrid_liquid_sim runs): ICCG on the device; the compressed-system flag
            // selects the same linear system without its identity rows in the reference and is solved uncompressed here
            mg = fromJet;
            if (this->useCompressedLinearSystem()) mg.linear_solver |= JETCU_SOLVER_COMPRESSED;
        } else if (!_allowSolverSubstitution)
            throw std::runtime_error(
                "jetcu: the pressure solver installed on this solver carries neither an MGPCG nor an ICCG linear system solver; "
                "the device path implements FdmMgpcgSolver3 and FdmIccgSolver3. Install one of them (or their Cuda* "
                "counterparts), or call setAllowSolverSubstitution(true) to use the constructor's MGPCG parameters");
        p.mgpcg = mg;
        if (_h && std::memcmp(&p, &_params, sizeof(p)) == 0) return;
        if (_h && std::memcmp(&p.grid, &_params.grid, sizeof(p.grid)) == 0) {
            _ctx->check(jetcu_hybrid_update_params(_h, &p));   // knobs changed between steps: device state stays
            _params = p;
            return;
        }
        if (_h) {   // the grid itself changed (resizeGrid): the live particles move to the new device solver through the host
            if (!_hostDirty) syncToHost();
            releaseDevice();
        }
        _ctx->check(jetcu_hybrid_create(_ctx->get(), &p, &_h));
        _params = p;
        _hostDirty = true;
    }

    void uploadIfNeeded() {
        auto particles = this->particleSystemData();
        const size_t n = particles->numberOfParticles();
        if (!_hostDirty && n > _deviceCount && _deviceCount > 0) {
            // a host emitter appended particles (ParticleSystemData3::addParticles) while the device holds the live copy of
            // the old ones: only the tail goes up, the device-resident state (APIC affine vectors included) stays
            _ctx->check(jetcu_hybrid_append_particles(
                _h, n - _deviceCount, reinterpret_cast<const double*>(particles->positions().data() + _deviceCount),
                reinterpret_cast<const double*>(particles->velocities().data() + _deviceCount)));
            _deviceCount = n;
            return;
        }
        if (_hostDirty || n != _deviceCount) {
            // host state is authoritative: (re)upload everything, including the velocity grid
            _ctx->check(jetcu_hybrid_set_particles(_h, n, reinterpret_cast<const double*>(particles->positions().data()),
                                                   reinterpret_cast<const double*>(particles->velocities().data())));
            auto vel = this->velocity();
            _ctx->check(jetcu_hybrid_set_velocity(_h, vel->uConstAccessor().data(), vel->vConstAccessor().data(),
                                                  vel->wConstAccessor().data()));
            _deviceCount = n;
            _hostDirty = false;
        }
    }

    //! Forwards the Jet collider to the device.  Accepted: a RigidBodyCollider3 whose surface is a Box3, Sphere3,
    //! Cylinder3 or Plane3 (normal-flipped or not), or an ImplicitSurfaceSet3 of up to JETCU_COLLIDER_MAX_SHAPES such
    //! surfaces added with addExplicitSurface — what the hybrid_liquid_sim examples build
    //! (src/examples/hybrid_liquid_sim/main.cpp:253-278, 454-463) — all with identity transforms; anything else is an
    //! error, never a silent fallback.  Re-sent when anything changed: the application may alter the linear / angular
    //! velocity between steps.
    void syncCollider() {
        const jet::Collider3Ptr col = this->collider();
        if (col == nullptr) {
            if (_hasCollider) _ctx->check(jetcu_hybrid_set_collider_shapes(_h, nullptr, 0, 0, nullptr, nullptr, 0.0));
            _hasCollider = false;
            return;
        }
        auto rigid = std::dynamic_pointer_cast<jet::RigidBodyCollider3>(col);
        if (!rigid) throw std::runtime_error("jetcu: only RigidBodyCollider3 colliders run on the device");
        std::vector<jetcu_collider_shape> shapes;
        int asSet = 0;
        detail::collectShapes(rigid->surface(), &shapes, &asSet);
        const double lin[3] = {rigid->linearVelocity.x, rigid->linearVelocity.y, rigid->linearVelocity.z};
        const double ang[3] = {rigid->angularVelocity.x, rigid->angularVelocity.y, rigid->angularVelocity.z};
        const double fr = col->frictionCoefficient();
        std::vector<double> key = {double(asSet), lin[0], lin[1], lin[2], ang[0], ang[1], ang[2], fr};
        for (const auto& s : shapes) {
            const double part[8] = {double(s.kind), double(s.flipped), s.a[0], s.a[1], s.a[2], s.b[0], s.b[1], s.b[2]};
            key.insert(key.end(), part, part + 8);
        }
        if (!_hasCollider || key != _colliderKey) {
            _ctx->check(jetcu_hybrid_set_collider_shapes(_h, shapes.data(), (uint32_t)shapes.size(), asSet, lin, ang, fr));
            _colliderKey = key;
            _hasCollider = true;
        }
    }

    static double picBlend(jet::FlipSolver3* s) { return s->picBlendingFactor(); }
    static double picBlend(jet::PicSolver3*) { return 0.0; }

    ContextPtr _ctx;
    jetcu_mgpcg_params _mgpcg;
    jetcu_hybrid_params _params = jetcu_hybrid_params();
    bool _allowSolverSubstitution = false;
    jetcu_hybrid* _h = nullptr;
    size_t _deviceCount = 0;
    bool _hostDirty = true;
    HostMirror _mirror = HostMirror::kEveryStep;
    jetcu_step_stats _stats = jetcu_step_stats();
    bool _hasCollider = false;
    std::vector<double> _colliderKey;
};

// ------------------------------------------------------------------------------------------------ smoke (config 5)
//! GridSmokeSolver3 with the sub-step on the device: GridFluidSolver3::onAdvanceTimeStep
//! (src/jet/grid_fluid_solver3.cpp:179-210) runs as jetcu_smoke_step.  Strict drop-in: emitters (VolumeGridEmitter3 ...)
//! keep running on the host inside beginAdvanceTimeStep, so the fields go up before and come back after every sub-step.
//! Supported configuration = what the device path implements: analytic rigid colliders, viscosity / diffusion through Jet's default diffusion solver,
//! CubicSemiLagrangian3 (the default) or SemiLagrangian3 advection; anything else throws instead of silently computing
//! something different.
class CudaGridSmokeSolver3 : public jet::GridSmokeSolver3 {
 public:
    CudaGridSmokeSolver3(const jet::Size3& resolution, const jet::Vector3D& gridSpacing, const jet::Vector3D& gridOrigin,
                         const jetcu_mgpcg_params& mgpcg = makeMgpcgParams(100, 6, 5, 5, 20, 20, 1e-9, 1.5, true),
                         ContextPtr ctx = ContextPtr())
        : jet::GridSmokeSolver3(resolution, gridSpacing, gridOrigin), _ctx(ctx ? ctx : Context::shared()), _mgpcg(mgpcg) {}

    ~CudaGridSmokeSolver3() override {
        if (_h) jetcu_smoke_destroy(_h);
    }

    const jetcu_mgpcg_result& lastResult() const { return _last; }

    //! kEveryStep (default): strict drop-in, the fields go up before and come back after every sub-step.  kOnDemand: the
    //! fields stay on the device between sub-steps as long as no host-side Jet emitter is attached; syncToHost() fetches
    //! them, markHostDirty() announces host-side edits.
    void setHostMirror(HostMirror m) { _mirror = m; }
    void markHostDirty() { _hostDirty = true; }
    void syncToHost() {
        if (!_h) return;
        auto vel = gridSystemData()->velocity();
        _ctx->check(jetcu_smoke_get_fields(_h, smokeDensity()->dataAccessor().data(), temperature()->dataAccessor().data(),
                                           vel->uAccessor().data(), vel->vAccessor().data(), vel->wAccessor().data(), nullptr));
    }

    //! Device-side smoke source: the equivalent of a continuous jet::VolumeGridEmitter3 over `surface` with
    //! addStepFunctionTarget(smokeDensity(), dMin, dMax) and addStepFunctionTarget(temperature(), tMin, tMax) — what
    //! src/examples/smoke_sim attaches — fired on the GPU before every sub-step (jetcu_smoke_emit_step).  The mapper of a Jet
    //! emitter target is an opaque lambda, so this cannot be read off a Jet emitter object; a Jet emitter set with setEmitter()
    //! keeps running on the host.
    void addDeviceStepEmitter(const jet::Surface3Ptr& surface, double dMin, double dMax, double tMin, double tMax) {
        DeviceEmitter e;
        detail::collectShapes(surface, &e.shapes, &e.asSet);
        e.v[0] = dMin; e.v[1] = dMax; e.v[2] = tMin; e.v[3] = tMax;
        _deviceEmitters.push_back(e);
    }

 protected:
    void onAdvanceTimeStep(double timeIntervalInSeconds) override {
        auto grids = gridSystemData();
        const jet::Size3 res = grids->resolution();
        if (res.x == 0 || res.y == 0 || res.z == 0) return;
        // computeViscosity / computeDiffusion run on the device with what a default-constructed solver carries
        // (GridBackwardEulerDiffusionSolver3, Neumann, FdmIccgSolver3(100, kEpsilonD)); another diffusion solver is refused
        const bool diffusive = viscosityCoefficient() > jet::kEpsilonD || smokeDiffusionCoefficient() > jet::kEpsilonD ||
                               temperatureDiffusionCoefficient() > jet::kEpsilonD;
        if (diffusive && !std::dynamic_pointer_cast<jet::GridBackwardEulerDiffusionSolver3>(diffusionSolver()))
            throw std::runtime_error("jetcu: viscosity / diffusion on the device need Jet's GridBackwardEulerDiffusionSolver3 "
                                     "(the solver's default); another GridDiffusionSolver3 is installed");
        // exactly CubicSemiLagrangian3 or exactly SemiLagrangian3 (its base class); any other subclass samples differently
        if (advectionSolver() == nullptr || (typeid(*advectionSolver()) != typeid(jet::CubicSemiLagrangian3) &&
                                             typeid(*advectionSolver()) != typeid(jet::SemiLagrangian3)))
            throw std::runtime_error("jetcu: the device smoke solver implements SemiLagrangian3 / CubicSemiLagrangian3 advection");
        ensureDevice();
        _ctx->check(jetcu_smoke_set_diffusion(_h, std::max(viscosityCoefficient(), 0.0), std::max(smokeDiffusionCoefficient(), 0.0),
                                              std::max(temperatureDiffusionCoefficient(), 0.0)));
        syncCollider();
        // the host part of beginAdvanceTimeStep: Jet emitters write into the host grids (src/jet/grid_fluid_solver3.cpp:437-447)
        const bool hostEmitter = emitter() != nullptr;
        const bool resident = _mirror == HostMirror::kOnDemand && !hostEmitter;
        if (hostEmitter) {
            if (_mirror == HostMirror::kOnDemand && !_hostDirty && _uploadedOnce) syncToHost();   // the emitter reads the current fields
            emitter()->update(currentTimeInSeconds(), timeIntervalInSeconds);
        }
        if (!resident || _hostDirty || !_uploadedOnce) {
            auto vel = grids->velocity();
            _ctx->check(jetcu_smoke_set_fields(_h, smokeDensity()->constDataAccessor().data(), temperature()->constDataAccessor().data(),
                                               vel->uConstAccessor().data(), vel->vConstAccessor().data(), vel->wConstAccessor().data()));
            _hostDirty = false;
            _uploadedOnce = true;
        }
        for (const DeviceEmitter& e : _deviceEmitters)
            _ctx->check(jetcu_smoke_emit_step(_h, e.shapes.data(), (uint32_t)e.shapes.size(), e.asSet, e.v[0], e.v[1], e.v[2], e.v[3]));
        _ctx->check(jetcu_smoke_step(_h, timeIntervalInSeconds, &_last));
        if (_mirror == HostMirror::kEveryStep) syncToHost();
    }

    //! GridFluidSolver3::numberOfSubTimeSteps (src/jet/grid_fluid_solver3.cpp:212-217) on the device-resident velocity: in
    //! kOnDemand mode the host grid is stale, and the CFL number must come from the live field.
    unsigned int numberOfSubTimeSteps(double timeIntervalInSeconds) const override {
        if (!_h || !_uploadedOnce || _hostDirty || _mirror == HostMirror::kEveryStep)
            return jet::GridSmokeSolver3::numberOfSubTimeSteps(timeIntervalInSeconds);
        uint32_t n = 1;
        _ctx->check(jetcu_smoke_num_sub_steps(_h, timeIntervalInSeconds, &n));
        return n;
    }

 private:
    void ensureDevice() {
        jetcu_smoke_params p;
        jetcu_smoke_default_params(&p);
        auto grids = gridSystemData();
        const jet::Size3 res = grids->resolution();
        const jet::Vector3D h = grids->gridSpacing(), o = grids->origin();
        p.grid.nx = res.x; p.grid.ny = res.y; p.grid.nz = res.z;
        p.grid.hx = h.x; p.grid.hy = h.y; p.grid.hz = h.z;
        p.grid.ox = o.x; p.grid.oy = o.y; p.grid.oz = o.z;
        p.closed_domain_boundary_flag = closedDomainBoundaryFlag();
        const jet::Vector3D g = gravity();
        p.gravity[0] = g.x; p.gravity[1] = g.y; p.gravity[2] = g.z;
        p.max_cfl = maxCfl();
        p.buoyancy_density_factor = buoyancySmokeDensityFactor();
        p.buoyancy_temperature_factor = buoyancyTemperatureFactor();
        p.density_decay = smokeDecayFactor();
        p.temperature_decay = smokeTemperatureDecayFactor();
        p.advection = typeid(*advectionSolver()) == typeid(jet::SemiLagrangian3) ? JETCU_ADVECT_LINEAR : JETCU_ADVECT_CUBIC;
        jetcu_mgpcg_params fromJet;
        auto frac = std::dynamic_pointer_cast<jet::GridFractionalSinglePhasePressureSolver3>(pressureSolver());
        if (frac && detail::mgpcgParamsOf(frac->linearSystemSolver(), &fromJet)) _mgpcg = fromJet;
        p.mgpcg = _mgpcg;
        if (_h && std::memcmp(&p, &_params, sizeof(p)) == 0) return;   // knobs unchanged since the last step
        if (_h) {   // knobs changed: the live fields move to the new device solver through the host
            if (_uploadedOnce && !_hostDirty) syncToHost();
            jetcu_smoke_destroy(_h);
        }
        _h = nullptr;
        _uploadedOnce = false;
        _ctx->check(jetcu_smoke_create(_ctx->get(), &p, &_h));
        _params = p;
    }

    //! RigidBodyCollider3 over Box3 / Sphere3 / Cylinder3 / Plane3 or an ImplicitSurfaceSet3 of them (identity transforms), as in
    //! the liquid adapter; re-sent when anything about it changed.
    void syncCollider() {
        const jet::Collider3Ptr col = collider();
        std::vector<double> key;
        std::vector<jetcu_collider_shape> shapes;
        int asSet = 0;
        double lin[3] = {0, 0, 0}, ang[3] = {0, 0, 0}, fr = 0.0;
        if (col != nullptr) {
            auto rigid = std::dynamic_pointer_cast<jet::RigidBodyCollider3>(col);
            if (!rigid) throw std::runtime_error("jetcu: only RigidBodyCollider3 colliders run on the device");
            detail::collectShapes(rigid->surface(), &shapes, &asSet);
            lin[0] = rigid->linearVelocity.x; lin[1] = rigid->linearVelocity.y; lin[2] = rigid->linearVelocity.z;
            ang[0] = rigid->angularVelocity.x; ang[1] = rigid->angularVelocity.y; ang[2] = rigid->angularVelocity.z;
            fr = col->frictionCoefficient();
            key = {double(asSet), lin[0], lin[1], lin[2], ang[0], ang[1], ang[2], fr};
            for (const auto& s : shapes) {
                const double part[8] = {double(s.kind), double(s.flipped), s.a[0], s.a[1], s.a[2], s.b[0], s.b[1], s.b[2]};
                key.insert(key.end(), part, part + 8);
            }
        }
        if (key != _colliderKey || _colliderHandle != _h) {
            _ctx->check(jetcu_smoke_set_collider_shapes(_h, shapes.data(), (uint32_t)shapes.size(), asSet, lin, ang, fr));
            _colliderKey = key;
            _colliderHandle = _h;
        }
    }

    struct DeviceEmitter {
        std::vector<jetcu_collider_shape> shapes;
        int asSet = 0;
        double v[4] = {0, 1, 0, 1};
    };

    ContextPtr _ctx;
    jetcu_mgpcg_params _mgpcg;
    jetcu_smoke* _h = nullptr;
    HostMirror _mirror = HostMirror::kEveryStep;
    bool _hostDirty = true, _uploadedOnce = false;
    std::vector<DeviceEmitter> _deviceEmitters;
    std::vector<double> _colliderKey;
    jetcu_smoke* _colliderHandle = nullptr;   // the device solver the key was sent to (ensureDevice may recreate it)
    jetcu_smoke_params _params = jetcu_smoke_params();
    jetcu_mgpcg_result _last = jetcu_mgpcg_result();
};
typedef std::shared_ptr<CudaGridSmokeSolver3> CudaGridSmokeSolver3Ptr;

typedef CudaHybridSolver3<jet::PicSolver3, JETCU_PIC> CudaPicSolver3;
typedef CudaHybridSolver3<jet::FlipSolver3, JETCU_FLIP> CudaFlipSolver3;
typedef CudaHybridSolver3<jet::ApicSolver3, JETCU_APIC> CudaApicSolver3;

}  // namespace jetcu

#endif  // JETCU_JET_ADAPTERS_H_
