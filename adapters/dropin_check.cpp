// dropin_check.cpp — a Jet application (public Jet API only, no `#define private public`) that runs the same FLIP /
// APIC / PIC dam-break scene through
//     stock   jet::FlipSolver3 + GridFractionalSinglePhasePressureSolver3 + FdmMgpcgSolver3   (the CPU reference)
//     seam1   the same with jetcu::CudaFdmMgpcgSolver3 installed via setLinearSystemSolver
//     seam2   the same with jetcu::CudaGridFractionalSinglePhasePressureSolver3 installed via setPressureSolver
//     seam3   jetcu::CudaFlipSolver3 (device-resident sub-step)
// using Animation::update(frame) exactly as src/examples/hybrid_liquid_sim/main.cpp does, and prints one JSON line
// with wall times and deviations from the stock run.  Built by adapters/Makefile against the unmodified reference
// (oracle/_ref/libjet_ref.a) and libjetcu.so;  tests/test_gpu_dropin.py runs it on the GPU box.
//
//     dropin_check [--n 32] [--frames 3] [--fps 60] [--levels 4] [--mode flip|apic|pic] [--skip-seam12] [--natural-ordering]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "jetcu_jet_adapters.h"

using namespace jet;

namespace {

struct Options {
    size_t n = 32;
    int frames = 3;
    double fps = 60.0;
    size_t levels = 4;
    std::string mode = "flip";
    bool skipSeam12 = false;
    bool redBlack = true;
};

double nowS() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

template <class Solver>
void setupScene(const std::shared_ptr<Solver>& solver, const Options& opt) {
    // cubic dam break, box (0,0,0)-(0.4,0.6,1.0), lattice spacing h/2, one shot, no jitter
    const double h = 1.0 / static_cast<double>(opt.n);
    auto box = std::make_shared<Box3>(Vector3D(0, 0, 0), Vector3D(0.4, 0.6, 1.0));
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
}

template <class Solver>
std::shared_ptr<Solver> makeSolver(const Options& opt) {
    const double h = 1.0 / static_cast<double>(opt.n);
    return std::make_shared<Solver>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D());
}

struct RunResult {
    std::vector<Vector3D> positions;
    std::vector<double> pressure;
    double seconds = 0.0;
    unsigned int lastIterations = 0;
};

template <class Solver>
void runFrames(const std::shared_ptr<Solver>& solver, const Options& opt, RunResult* out) {
    double t0 = nowS();
    for (Frame frame(0, 1.0 / opt.fps); frame.index < opt.frames; ++frame) solver->update(frame);
    out->seconds = nowS() - t0;
    auto pos = solver->particleSystemData()->positions();
    out->positions.assign(pos.data(), pos.data() + pos.size());
}

double maxPositionDeviation(const RunResult& a, const RunResult& b) {
    if (a.positions.size() != b.positions.size()) return INFINITY;
    double m = 0.0;
    for (size_t i = 0; i < a.positions.size(); ++i) {
        Vector3D d = a.positions[i] - b.positions[i];
        m = std::max(m, std::max(std::fabs(d.x), std::max(std::fabs(d.y), std::fabs(d.z))));
    }
    return m;
}

double relLinf(const std::vector<double>& a, const std::vector<double>& b) {
    if (a.size() != b.size() || a.empty()) return INFINITY;
    double m = 0.0, s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) {
        m = std::max(m, std::fabs(a[i] - b[i]));
        s = std::max(s, std::fabs(b[i]));
    }
    return m / std::max(s, 1e-300);
}

std::shared_ptr<FdmMgpcgSolver3> stockMgpcg(const Options& opt) {
    return std::make_shared<FdmMgpcgSolver3>(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, opt.redBlack);
}

template <class Stock, class Cuda>
int runMode(const Options& opt) {
    RunResult stock, seam1, seam2, seam3;

    {  // ---- stock reference
        auto solver = makeSolver<Stock>(opt);
        setupScene(solver, opt);
        auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
        auto lin = stockMgpcg(opt);
        pressure->setLinearSystemSolver(lin);
        solver->setPressureSolver(pressure);
        runFrames(solver, opt, &stock);
        stock.pressure.assign(pressure->pressure().data(), pressure->pressure().data() + opt.n * opt.n * opt.n);
        stock.lastIterations = lin->lastNumberOfIterations();
    }
    if (!opt.skipSeam12) {
        {  // ---- seam 1: GPU linear solver behind FdmMgSolver3
            auto solver = makeSolver<Stock>(opt);
            setupScene(solver, opt);
            auto pressure = std::make_shared<GridFractionalSinglePhasePressureSolver3>();
            auto lin = std::make_shared<jetcu::CudaFdmMgpcgSolver3>(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, opt.redBlack);
            pressure->setLinearSystemSolver(lin);
            solver->setPressureSolver(pressure);
            runFrames(solver, opt, &seam1);
            seam1.pressure.assign(pressure->pressure().data(), pressure->pressure().data() + opt.n * opt.n * opt.n);
            seam1.lastIterations = lin->lastNumberOfIterations();
        }
        {  // ---- seam 2: GPU pressure solver behind GridPressureSolver3
            auto solver = makeSolver<Stock>(opt);
            setupScene(solver, opt);
            auto pressure = std::make_shared<jetcu::CudaGridFractionalSinglePhasePressureSolver3>(
                jetcu::makeMgpcgParams(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, opt.redBlack));
            solver->setPressureSolver(pressure);
            runFrames(solver, opt, &seam2);
            seam2.pressure.assign(pressure->pressure().data(), pressure->pressure().data() + opt.n * opt.n * opt.n);
            seam2.lastIterations = pressure->lastResult().iterations;
        }
    }
    {  // ---- seam 3: device-resident solver subclass
        const double h = 1.0 / static_cast<double>(opt.n);
        auto solver = std::make_shared<Cuda>(Size3(opt.n, opt.n, opt.n), Vector3D(h, h, h), Vector3D(),
                                             jetcu::makeMgpcgParams(100, opt.levels, 5, 5, 20, 20, 1e-9, 1.5, opt.redBlack));
        setupScene(solver, opt);
        runFrames(solver, opt, &seam3);
        seam3.lastIterations = solver->lastStepStats().pcg_iterations;
    }

    std::printf("{\"mode\": \"%s\", \"n\": %zu, \"frames\": %d, \"particles\": %zu, \"stock_s\": %.4f, "
                "\"stock_last_iterations\": %u",
                opt.mode.c_str(), opt.n, opt.frames, stock.positions.size(), stock.seconds, stock.lastIterations);
    if (!opt.skipSeam12) {
        std::printf(", \"seam1_s\": %.4f, \"seam1_max_dx\": %.3e, \"seam1_pressure_rel\": %.3e, \"seam1_last_iterations\": %u",
                    seam1.seconds, maxPositionDeviation(seam1, stock), relLinf(seam1.pressure, stock.pressure),
                    seam1.lastIterations);
        std::printf(", \"seam2_s\": %.4f, \"seam2_max_dx\": %.3e, \"seam2_pressure_rel\": %.3e, \"seam2_last_iterations\": %u",
                    seam2.seconds, maxPositionDeviation(seam2, stock), relLinf(seam2.pressure, stock.pressure),
                    seam2.lastIterations);
    }
    std::printf(", \"seam3_s\": %.4f, \"seam3_max_dx\": %.3e, \"seam3_last_iterations\": %u}\n", seam3.seconds,
                maxPositionDeviation(seam3, stock), seam3.lastIterations);
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
        else if (a == "--fps") opt.fps = std::atof(next());
        else if (a == "--levels") opt.levels = std::strtoul(next(), nullptr, 10);
        else if (a == "--mode") opt.mode = next();
        else if (a == "--skip-seam12") opt.skipSeam12 = true;
        else if (a == "--natural-ordering") opt.redBlack = false;   // FdmMgpcgSolver3's constructor default (useRedBlackOrdering = false)
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    Logging::mute();
    try {
        if (opt.mode == "flip") return runMode<FlipSolver3, jetcu::CudaFlipSolver3>(opt);
        if (opt.mode == "apic") return runMode<ApicSolver3, jetcu::CudaApicSolver3>(opt);
        if (opt.mode == "pic") return runMode<PicSolver3, jetcu::CudaPicSolver3>(opt);
        std::fprintf(stderr, "unknown mode %s\n", opt.mode.c_str());
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "dropin_check: %s\n", e.what());
        return 1;
    }
}
