// smoke_sim_cuda.cpp — Jet's smoke_sim example 1 (src/examples/smoke_sim/main.cpp:189-235: rising smoke from a continuous box
// source under a sphere, cubic semi-Lagrangian advection) on jetcu::CudaGridSmokeSolver3: device-side source, collider forwarded
// from the Jet object, fields resident on the device between sub-steps.  Public Jet API + adapters/jetcu_jet_adapters.h only;
// the scene is re-described on a cubic grid (see hybrid_liquid_sim_cuda.cpp for why).
//
//     smoke_sim_cuda [-r 64] [-f 30] [--linear] [--compare]
//
// --linear installs SemiLagrangian3 instead of the default CubicSemiLagrangian3 (example 2); --compare steps a stock
// GridSmokeSolver3 with Jet's own VolumeGridEmitter3 alongside and prints the largest density deviation per frame.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "jetcu_jet_adapters.h"

using namespace jet;

int main(int argc, char** argv) {
    size_t r = 64;
    int frames = 30;
    bool linear = false, compare = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-r" && i + 1 < argc) r = static_cast<size_t>(std::atoi(argv[++i]));
        else if (a == "-f" && i + 1 < argc) frames = std::atoi(argv[++i]);
        else if (a == "--linear") linear = true;
        else if (a == "--compare") compare = true;
        else { std::fprintf(stderr, "usage: %s [-r N] [-f frames] [--linear] [--compare]\n", argv[0]); return 2; }
    }
    Logging::mute();
    try {
        const double h = 1.0 / static_cast<double>(r);
        size_t levels = 1;
        for (size_t n = r; n % 2 == 0 && n > 8 && levels < 6; n /= 2) ++levels;
        auto source = Box3::builder().withLowerCorner({0.45, -1, 0.45}).withUpperCorner({0.55, 0.05, 0.55}).makeShared();
        auto sphere = Sphere3::builder().withCenter({0.5, 0.3, 0.5}).withRadius(0.075).makeShared();

        auto cuda = std::make_shared<jetcu::CudaGridSmokeSolver3>(Size3(r, r, r), Vector3D(h, h, h), Vector3D(),
                                                                  jetcu::makeMgpcgParams(100, levels, 5, 5, 20, 20, 1e-9, 1.5, true));
        if (linear) cuda->setAdvectionSolver(std::make_shared<SemiLagrangian3>());
        cuda->setCollider(RigidBodyCollider3::builder().withSurface(sphere).makeShared());
        cuda->addDeviceStepEmitter(source, 0, 1, 0, 1);
        cuda->setHostMirror(compare ? jetcu::HostMirror::kEveryStep : jetcu::HostMirror::kOnDemand);

        std::shared_ptr<GridSmokeSolver3> stock;
        if (compare) {
            stock = std::make_shared<GridSmokeSolver3>(Size3(r, r, r), Vector3D(h, h, h), Vector3D());
            auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
            pressure->setLinearSystemSolver(std::make_shared<FdmMgpcgSolver3>(100, levels, 5, 5, 20, 20, 1e-9, 1.5, true));
            stock->setPressureSolver(pressure);
            if (linear) stock->setAdvectionSolver(std::make_shared<SemiLagrangian3>());
            stock->setCollider(RigidBodyCollider3::builder().withSurface(sphere).makeShared());
            auto emitter = VolumeGridEmitter3::builder().withSourceRegion(source).withIsOneShot(false).makeShared();
            stock->setEmitter(emitter);
            emitter->addStepFunctionTarget(stock->smokeDensity(), 0, 1);
            emitter->addStepFunctionTarget(stock->temperature(), 0, 1);
        }
        for (Frame frame(0, 1.0 / 60.0); frame.index < frames; ++frame) {
            const auto t0 = std::chrono::steady_clock::now();
            cuda->update(frame);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            std::printf("frame %d: %.1f ms, %u PCG iterations", frame.index, ms, cuda->lastResult().iterations);
            if (stock) {
                stock->update(frame);
                const auto a = stock->smokeDensity()->constDataAccessor();
                const auto b = cuda->smokeDensity()->constDataAccessor();
                double d = 0.0;
                for (size_t i = 0; i < r * r * r; ++i) d = std::max(d, std::fabs(a.data()[i] - b.data()[i]));
                std::printf(", max |density - stock Jet| %.3e", d);
            }
            std::printf("\n");
        }
        if (!compare) cuda->syncToHost();
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "smoke_sim_cuda: %s\n", e.what());
        return 1;
    }
}
