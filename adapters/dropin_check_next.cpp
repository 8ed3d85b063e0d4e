// dropin_check_next.cpp — Jet-side check (public Jet API only) of the two adapter paths added after the three seams:
//     collider   jet::FlipSolver3 + RigidBodyCollider3(Box3) through Animation::update  vs  jetcu::CudaFlipSolver3 with the
//                same collider object (adapter forwards it as jetcu_hybrid_set_collider_shapes)
//     pillars    the same with the collider of hybrid_liquid_sim example 3: an ImplicitSurfaceSet3 of three Cylinder3
//     tank       the same inside a normal-flipped Sphere3 (example 6)
//     compressed stock Jet only, no GPU: FlipSolver3 + MGPCG with setUseCompressedLinearSystem(true) vs (false) — with a
//                multigrid solver installed the reference ignores the flag, which is why the adapters may accept it
//     emit       the water drop of examples 1 / 2 (half-space + sphere): stock VolumeParticleEmitter3 set on the host  vs
//                CudaFlipSolver3::emitOnDevice (lattice generated and tested on the GPU), then the same frames
//     smoke      jet::GridSmokeSolver3 (plume from rest in a closed box)               vs  jetcu::CudaGridSmokeSolver3
//     smoke_collider  the same with smoke_sim example 1's continuous box source and sphere collider
//     smoke_device    ... with the source as a device-side emitter and the fields resident on the device (kOnDemand)
// Prints one JSON line with the deviations from the stock run.  Built by adapters/Makefile next to dropin_check;
// tests/test_gpu_zz_dropin_next.py runs it (first GPU run pending, see that file).
//
//     dropin_check_next [--n 16] [--frames 3] [--levels 3] [--what collider|pillars|tank|emit|compressed|smoke|smoke_collider|smoke_device]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "jetcu_jet_adapters.h"

using namespace jet;

namespace {

struct Options {
    size_t n = 16;
    int frames = 3;
    size_t levels = 3;
    std::string what = "collider";
};

template <class Solver>
void setupLiquid(const std::shared_ptr<Solver>& solver, const Options& opt) {
    const double h = 1.0 / static_cast<double>(opt.n);
    auto box = std::make_shared<Box3>(Vector3D(0, 0, 0), Vector3D(0.5, 0.7, 1.0));
    auto emitter = VolumeParticleEmitter3::builder()
                       .withSurface(box)
                       .withMaxRegion(solver->gridSystemData()->boundingBox())
                       .withSpacing(0.5 * h)
                       .withIsOneShot(true)
                       .withJitter(0.0)
                       .makeShared();
    emitter->setPointGenerator(std::make_shared<GridPointGenerator3>());
    solver->setParticleEmitter(emitter);
    solver->setUseCompressedLinearSystem(false);
    Surface3Ptr obstacle;
    if (opt.what == "pillars") {
        obstacle = ImplicitSurfaceSet3::builder()
                       .withExplicitSurfaces({std::make_shared<Cylinder3>(Vector3D(0.65, 0.25, 0.2), 0.08, 0.5),
                                              std::make_shared<Cylinder3>(Vector3D(0.7, 0.25, 0.5), 0.08, 0.5),
                                              std::make_shared<Cylinder3>(Vector3D(0.65, 0.25, 0.8), 0.08, 0.5)})
                       .makeShared();
    } else if (opt.what == "tank") {
        obstacle = Sphere3::builder().withCenter({0.5, 0.5, 0.5}).withRadius(0.48).withIsNormalFlipped(true).makeShared();
    } else {
        obstacle = std::make_shared<Box3>(Vector3D(0.6, 0.0, 0.3), Vector3D(0.85, 0.35, 0.7));
    }
    auto collider = std::make_shared<RigidBodyCollider3>(obstacle);
    collider->setFrictionCoefficient(0.2);
    solver->setCollider(collider);
}

int runCollider(const Options& opt) {
    const double h = 1.0 / static_cast<double>(opt.n);
    auto stock = std::make_shared<FlipSolver3>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D());
    {
        auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
        pressure->setLinearSystemSolver(std::make_shared<FdmMgpcgSolver3>(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, true));
        stock->setPressureSolver(pressure);
    }
    setupLiquid(stock, opt);
    auto cuda = std::make_shared<jetcu::CudaFlipSolver3>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D(),
                                                         jetcu::makeMgpcgParams(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, true));
    setupLiquid(cuda, opt);
    for (Frame frame(0, 1.0 / 60.0); frame.index < opt.frames; ++frame) {
        stock->update(frame);
        cuda->update(frame);
    }
    auto ps = stock->particleSystemData();
    auto pc = cuda->particleSystemData();
    double maxDx = 0.0;
    const size_t n = ps->numberOfParticles();
    if (pc->numberOfParticles() != n) { std::fprintf(stderr, "particle counts differ\n"); return 1; }
    for (size_t i = 0; i < n; ++i) {
        const Vector3D d = ps->positions()[i] - pc->positions()[i];
        maxDx = std::max(maxDx, std::max(std::fabs(d.x), std::max(std::fabs(d.y), std::fabs(d.z))));
    }
    std::printf("{\"what\": \"%s\", \"n\": %zu, \"frames\": %d, \"particles\": %zu, \"max_dx\": %.3e}\n", opt.what.c_str(), opt.n, opt.frames, n, maxDx);
    return 0;
}

int runCompressed(const Options& opt) {
    const double h = 1.0 / static_cast<double>(opt.n);
    std::shared_ptr<FlipSolver3> solvers[2];
    for (int q = 0; q < 2; ++q) {
        solvers[q] = std::make_shared<FlipSolver3>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D());
        auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
        pressure->setLinearSystemSolver(std::make_shared<FdmMgpcgSolver3>(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, true));
        solvers[q]->setPressureSolver(pressure);
        Options scene = opt;
        scene.what = "collider";
        setupLiquid(solvers[q], scene);                       // sets the flag to false ...
        solvers[q]->setUseCompressedLinearSystem(q == 1);     // ... and here it is switched on for the second solver
    }
    for (Frame frame(0, 1.0 / 60.0); frame.index < opt.frames; ++frame)
        for (auto& s : solvers) s->update(frame);
    auto p0 = solvers[0]->particleSystemData();
    auto p1 = solvers[1]->particleSystemData();
    const size_t n = p0->numberOfParticles();
    if (p1->numberOfParticles() != n) { std::fprintf(stderr, "particle counts differ\n"); return 1; }
    size_t different = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::memcmp(&p0->positions()[i], &p1->positions()[i], sizeof(Vector3D)) != 0 ||
            std::memcmp(&p0->velocities()[i], &p1->velocities()[i], sizeof(Vector3D)) != 0) ++different;
    std::printf("{\"what\": \"compressed\", \"n\": %zu, \"frames\": %d, \"particles\": %zu, \"different\": %zu}\n", opt.n, opt.frames, n, different);
    return 0;
}

int runEmit(const Options& opt) {
    const double h = 1.0 / static_cast<double>(opt.n);
    auto stock = std::make_shared<FlipSolver3>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D());
    {
        auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
        pressure->setLinearSystemSolver(std::make_shared<FdmMgpcgSolver3>(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, true));
        stock->setPressureSolver(pressure);
    }
    stock->setUseCompressedLinearSystem(false);
    auto cuda = std::make_shared<jetcu::CudaFlipSolver3>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D(),
                                                         jetcu::makeMgpcgParams(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, true));
    const BoundingBox3D domain = stock->gridSystemData()->boundingBox();
    auto plane = Plane3::builder().withNormal({0, 1, 0}).withPoint({0, 0.25 * domain.height(), 0}).makeShared();
    auto sphere = Sphere3::builder().withCenter(domain.midPoint()).withRadius(0.15 * domain.width()).makeShared();
    auto e1 = VolumeParticleEmitter3::builder().withSurface(plane).withSpacing(0.5 * h).withMaxRegion(domain).withIsOneShot(true).makeShared();
    e1->setPointGenerator(std::make_shared<GridPointGenerator3>());
    auto e2 = VolumeParticleEmitter3::builder().withSurface(sphere).withSpacing(0.5 * h).withMaxRegion(domain).withIsOneShot(true).makeShared();
    e2->setPointGenerator(std::make_shared<GridPointGenerator3>());
    stock->setParticleEmitter(ParticleEmitterSet3::builder().withEmitters({e1, e2}).makeShared());
    const size_t emitted = cuda->emitOnDevice(plane, 0.5 * h) + cuda->emitOnDevice(sphere, 0.5 * h);
    for (Frame frame(0, 1.0 / 60.0); frame.index < opt.frames; ++frame) {
        stock->update(frame);
        cuda->update(frame);
    }
    auto ps = stock->particleSystemData();
    auto pc = cuda->particleSystemData();
    const size_t n = ps->numberOfParticles();
    if (pc->numberOfParticles() != n || emitted != n) { std::fprintf(stderr, "particle counts differ: %zu %zu %zu\n", n, pc->numberOfParticles(), emitted); return 1; }
    double maxDx = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const Vector3D d = ps->positions()[i] - pc->positions()[i];
        maxDx = std::max(maxDx, std::max(std::fabs(d.x), std::max(std::fabs(d.y), std::fabs(d.z))));
    }
    std::printf("{\"what\": \"emit\", \"n\": %zu, \"frames\": %d, \"particles\": %zu, \"max_dx\": %.3e}\n", opt.n, opt.frames, n, maxDx);
    return 0;
}

// smoke_device: the jetcu solver gets the source as a device-side emitter and keeps its fields on the device between sub-steps
inline bool attachDeviceEmitter(const std::shared_ptr<GridSmokeSolver3>&, const Surface3Ptr&, const Options&) { return false; }
inline bool attachDeviceEmitter(const std::shared_ptr<jetcu::CudaGridSmokeSolver3>& solver, const Surface3Ptr& source, const Options& opt) {
    if (opt.what != "smoke_device") return false;
    solver->addDeviceStepEmitter(source, 0, 1, 0, 1);
    solver->setHostMirror(jetcu::HostMirror::kOnDemand);
    return true;
}

template <class Solver>
void setupSmoke(const std::shared_ptr<Solver>& solver, const Options& opt) {
    solver->setUseCompressedLinearSystem(false);
    auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
    pressure->setLinearSystemSolver(std::make_shared<FdmMgpcgSolver3>(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, true));
    solver->setPressureSolver(pressure);
    auto den = solver->smokeDensity();
    auto temp = solver->temperature();
    auto pos = den->dataPosition();
    den->forEachDataPointIndex([&](size_t i, size_t j, size_t k) {
        const Vector3D p = pos(i, j, k);
        const double r = (p - Vector3D(0.5, 0.25, 0.45)).length();
        const double v = clamp(1.0 - r / 0.2, 0.0, 1.0);
        (*den)(i, j, k) = v;
        (*temp)(i, j, k) = 2.0 * v + 0.1;
    });
    if (opt.what == "smoke_collider" || opt.what == "smoke_device") {   // src/examples/smoke_sim example 1: a continuous box source under a sphere
        auto source = Box3::builder().withLowerCorner({0.45, -1, 0.45}).withUpperCorner({0.55, 0.05, 0.55}).makeShared();
        if (!attachDeviceEmitter(solver, source, opt)) {
            auto emitter = VolumeGridEmitter3::builder().withSourceRegion(source).withIsOneShot(false).makeShared();
            solver->setEmitter(emitter);
            emitter->addStepFunctionTarget(solver->smokeDensity(), 0, 1);
            emitter->addStepFunctionTarget(solver->temperature(), 0, 1);
        }
        auto sphere = Sphere3::builder().withCenter({0.5, 0.3, 0.5}).withRadius(0.075).makeShared();
        solver->setCollider(RigidBodyCollider3::builder().withSurface(sphere).makeShared());
    }
}

double fieldDiff(const ConstArrayAccessor3<double>& a, const ConstArrayAccessor3<double>& b, double* scale) {
    double d = 0.0, s = 0.0;
    const size_t n = a.size().x * a.size().y * a.size().z;
    for (size_t i = 0; i < n; ++i) { d = std::max(d, std::fabs(a.data()[i] - b.data()[i])); s = std::max(s, std::fabs(a.data()[i])); }
    *scale = std::max(*scale, s);
    return d;
}

int runSmoke(const Options& opt) {
    const double h = 1.0 / static_cast<double>(opt.n);
    auto stock = std::make_shared<GridSmokeSolver3>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D());
    auto cuda = std::make_shared<jetcu::CudaGridSmokeSolver3>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D());
    setupSmoke(stock, opt);
    setupSmoke(cuda, opt);
    for (Frame frame(0, 1.0 / 30.0); frame.index < opt.frames; ++frame) {
        stock->update(frame);
        cuda->update(frame);
    }
    cuda->syncToHost();   // kOnDemand (smoke_device): the fields lived on the device until now
    double vs = 0.0, ds = 0.0;
    auto sv = stock->velocity();
    auto cv = cuda->velocity();
    double dv = fieldDiff(sv->uConstAccessor(), cv->uConstAccessor(), &vs);
    dv = std::max(dv, fieldDiff(sv->vConstAccessor(), cv->vConstAccessor(), &vs));
    dv = std::max(dv, fieldDiff(sv->wConstAccessor(), cv->wConstAccessor(), &vs));
    const double dd = fieldDiff(stock->smokeDensity()->constDataAccessor(), cuda->smokeDensity()->constDataAccessor(), &ds);
    std::printf("{\"what\": \"smoke\", \"n\": %zu, \"frames\": %d, \"velocity_rel\": %.3e, \"density_rel\": %.3e, \"velocity_scale\": %.3e}\n",
                opt.n, opt.frames, dv / std::max(vs, 1e-300), dd / std::max(ds, 1e-300), vs);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    Options opt;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : "0"; };
        if (a == "--n") opt.n = std::strtoul(next(), nullptr, 10);
        else if (a == "--frames") opt.frames = std::atoi(next());
        else if (a == "--levels") opt.levels = std::strtoul(next(), nullptr, 10);
        else if (a == "--what") opt.what = next();
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    Logging::mute();
    try {
        if (opt.what == "collider" || opt.what == "pillars" || opt.what == "tank") return runCollider(opt);
        if (opt.what == "compressed") return runCompressed(opt);
        if (opt.what == "emit") return runEmit(opt);
        if (opt.what == "smoke" || opt.what == "smoke_collider" || opt.what == "smoke_device") return runSmoke(opt);
        std::fprintf(stderr, "unknown check %s\n", opt.what.c_str());
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "dropin_check_next: %s\n", e.what());
        return 1;
    }
}
