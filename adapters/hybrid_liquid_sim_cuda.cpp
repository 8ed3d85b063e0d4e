// hybrid_liquid_sim_cuda.cpp — what switching a Jet application over looks like: the scenes of Jet's own hybrid_liquid_sim
// example (src/examples/hybrid_liquid_sim/main.cpp) with jetcu's solver classes in place of Jet's.  Public Jet API plus
// adapters/jetcu_jet_adapters.h only; nothing of the example's source is reused — the scenes are re-described here on CUBIC
// grids so that --compare has a well-defined stock run next to it: on ny != nz the reference's own multi-level V-cycle reads
// out of bounds (FdmMgUtils3::correct clamps k against n.y, src/jet/fdm_mg_linear_system3.cpp:159), which is why Jet's example
// falls back to ICCG on its 1 x 2 x 1 domains.  (jetcu itself accepts such shapes, clamping against n.z.)
//
//     hybrid_liquid_sim_cuda [-e 1|3|6] [-r 64] [-f 10] [-o out_dir] [-m pos|xyz|none] [--host-emit] [--compare]
//
//   -e 1  water drop, FLIP : half-space + sphere emitters                              (main.cpp:95-151)
//   -e 3  dam break, FLIP  : two-box emitter set, three-pillar ImplicitSurfaceSet3 collider  (main.cpp:208-277)
//   -e 6  tank, APIC       : fluid inside a normal-flipped Sphere3 collider                  (main.cpp:445-495)
// Emission runs on the device (CudaHybridSolver3::emitOnDevice) unless --host-emit attaches Jet's VolumeParticleEmitter3;
// frames are written from the device-resident particle set (saveParticles, byte-identical to saveParticleAsPos / Xyz).
// --compare steps a stock Jet solver next to it and prints the largest particle deviation per frame.
#include <algorithm>
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
    int example = 3;
    size_t r = 64;
    int frames = 10;
    std::string out = ".";
    std::string format = "none";
    bool hostEmit = false, compare = false;
    bool iccg = false;   // the examples' own solver configuration: default FdmIccgSolver3 + setUseCompressedLinearSystem(true)
    double fps = 60.0;
};

struct Scene {
    std::vector<Surface3Ptr> sources;  // one emitter per entry, fired in order (ParticleEmitterSet3)
    Surface3Ptr collider;              // may be null
};

Scene makeScene(int example) {
    Scene s;
    if (example == 1) {
        s.sources.push_back(Plane3::builder().withNormal({0, 1, 0}).withPoint({0, 0.25, 0}).makeShared());
        s.sources.push_back(Sphere3::builder().withCenter({0.5, 0.6, 0.5}).withRadius(0.15).makeShared());
    } else if (example == 3) {
        s.sources.push_back(ImplicitSurfaceSet3::builder()
                                .withExplicitSurfaces({std::make_shared<Box3>(Vector3D(0, 0, 0), Vector3D(0.2 + 0.001, 0.5 + 0.001, 0.75 + 0.001)),
                                                       std::make_shared<Box3>(Vector3D(0.85 - 0.001, 0, 0.25 - 0.001), Vector3D(1.2, 0.5 + 0.001, 1.2))})
                                .makeShared());
        s.collider = ImplicitSurfaceSet3::builder()
                         .withExplicitSurfaces({std::make_shared<Cylinder3>(Vector3D(0.35, 0.25, 0.25), 0.04, 0.5),
                                                std::make_shared<Cylinder3>(Vector3D(0.5, 0.25, 0.5), 0.04, 0.5),
                                                std::make_shared<Cylinder3>(Vector3D(0.65, 0.25, 0.75), 0.04, 0.5)})
                         .makeShared();
    } else {
        s.sources.push_back(std::make_shared<Box3>(Vector3D(0.2, 0.2, 0.2), Vector3D(0.8, 0.5, 0.8)));
        s.collider = Sphere3::builder().withCenter({0.5, 0.5, 0.5}).withRadius(0.4).withIsNormalFlipped(true).makeShared();
    }
    return s;
}

template <class Solver>
void attachHostEmitters(const std::shared_ptr<Solver>& solver, const Scene& scene, double spacing) {
    std::vector<ParticleEmitter3Ptr> emitters;
    for (const auto& src : scene.sources) {
        auto e = VolumeParticleEmitter3::builder()
                     .withSurface(src)
                     .withSpacing(spacing)
                     .withMaxRegion(solver->gridSystemData()->boundingBox())
                     .withIsOneShot(true)
                     .makeShared();
        e->setPointGenerator(std::make_shared<GridPointGenerator3>());
        emitters.push_back(e);
    }
    solver->setParticleEmitter(ParticleEmitterSet3::builder().withEmitters(emitters).makeShared());
}

template <class Solver>
void attachCollider(const std::shared_ptr<Solver>& solver, const Scene& scene) {
    if (scene.collider) solver->setCollider(RigidBodyCollider3::builder().withSurface(scene.collider).makeShared());
}

template <class Stock, class Cuda>
int run(const Options& opt) {
    const double h = 1.0 / static_cast<double>(opt.r);
    const Size3 res(opt.r, opt.r, opt.r);
    const Scene scene = makeScene(opt.example);
    size_t levels = 1;
    for (size_t n = opt.r; n % 2 == 0 && n > 8 && levels < 6; n /= 2) ++levels;
    auto cuda = std::make_shared<Cuda>(res, Vector3D(h, h, h), Vector3D(), jetcu::makeMgpcgParams(100, levels, 5, 5, 20, 20, 1e-9, 1.5, true));
    if (opt.iccg) {   // src/examples/hybrid_liquid_sim/main.cpp:219: a default-constructed pressure solver (ICCG) + compressed system
        cuda->setPressureSolver(std::make_shared<GridFractionalSinglePhasePressureSolver3>());
        cuda->setUseCompressedLinearSystem(true);
    }
    cuda->setHostMirror(opt.compare ? jetcu::HostMirror::kEveryStep : jetcu::HostMirror::kOnDemand);
    attachCollider(cuda, scene);
    size_t emitted = 0;
    if (opt.hostEmit) attachHostEmitters(cuda, scene, 0.5 * h);
    else for (const auto& src : scene.sources) emitted += cuda->emitOnDevice(src, 0.5 * h);
    std::shared_ptr<Stock> stock;
    if (opt.compare) {
        stock = std::make_shared<Stock>(res, Vector3D(h, h, h), Vector3D());
        auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
        if (!opt.iccg) pressure->setLinearSystemSolver(std::make_shared<FdmMgpcgSolver3>(100, levels, 5, 5, 20, 20, 1e-9, 1.5, true));
        stock->setPressureSolver(pressure);
        stock->setUseCompressedLinearSystem(opt.iccg);
        attachCollider(stock, scene);
        attachHostEmitters(stock, scene, 0.5 * h);
    }
    Logging::mute();  // stock Jet logs every stage at INFO level to stdout; this program's frame lines are its output
    std::printf("example %d at %zu^3, %zu MG levels, %s emission%s\n", opt.example, opt.r, levels, opt.hostEmit ? "host" : "device",
                opt.hostEmit ? "" : (" (" + std::to_string(emitted) + " particles)").c_str());
    for (Frame frame(0, 1.0 / opt.fps); frame.index < opt.frames; ++frame) {
        const auto t0 = std::chrono::steady_clock::now();
        cuda->update(frame);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (opt.format != "none") {
            char name[64];
            std::snprintf(name, sizeof(name), "frame_%06d.%s", frame.index, opt.format.c_str());
            cuda->saveParticles(opt.out + "/" + name, opt.format == "pos" ? JETCU_FRAME_POS : JETCU_FRAME_XYZ);
        }
        std::printf("frame %d: %.1f ms, %u PCG iterations", frame.index, ms, cuda->lastStepStats().pcg_iterations);
        if (stock) {
            stock->update(frame);
            auto ps = stock->particleSystemData();
            auto pc = cuda->particleSystemData();
            double maxDx = -1.0;
            if (ps->numberOfParticles() == pc->numberOfParticles()) {
                maxDx = 0.0;
                for (size_t i = 0; i < ps->numberOfParticles(); ++i) {
                    const Vector3D d = ps->positions()[i] - pc->positions()[i];
                    maxDx = std::max(maxDx, std::max(std::fabs(d.x), std::max(std::fabs(d.y), std::fabs(d.z))));
                }
            }
            std::printf(", %zu / %zu particles, max |dx| vs stock Jet %.3e", pc->numberOfParticles(), ps->numberOfParticles(), maxDx);
        }
        std::printf("\n");
    }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    Options opt;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
        if (a == "-e") opt.example = std::atoi(next());
        else if (a == "-r") opt.r = static_cast<size_t>(std::atoi(next()));
        else if (a == "-f") opt.frames = std::atoi(next());
        else if (a == "-o") opt.out = next();
        else if (a == "-m") opt.format = next();
        else if (a == "--host-emit") opt.hostEmit = true;
        else if (a == "--compare") opt.compare = true;
        else if (a == "--iccg") opt.iccg = true;
        else { std::fprintf(stderr, "usage: %s [-e 1|3|6] [-r N] [-f frames] [-o dir] [-m pos|xyz|none] [--host-emit] [--compare] [--iccg]\n", argv[0]); return 2; }
    }
    if (opt.format != "pos" && opt.format != "xyz" && opt.format != "none") { std::fprintf(stderr, "unknown format %s\n", opt.format.c_str()); return 2; }
    try {
        if (opt.example == 1) return run<FlipSolver3, jetcu::CudaFlipSolver3>(opt);
        if (opt.example == 3) return run<FlipSolver3, jetcu::CudaFlipSolver3>(opt);
        if (opt.example == 6) return run<ApicSolver3, jetcu::CudaApicSolver3>(opt);
        std::fprintf(stderr, "examples: 1 (water drop), 3 (dam break with pillars), 6 (spherical tank)\n");
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "hybrid_liquid_sim_cuda: %s\n", e.what());
        return 1;
    }
}
