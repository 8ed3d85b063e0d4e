#!/usr/bin/env python
"""bench.py — sub-steps/s of the B200-native hybrid-liquid path (BASELINE.json metric: FLIP 256^3), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl jetcu|reference]
                    [--workload tank|slosh|dam] [--mode flip|apic|pic] [--n 256]

* own arm (default): jetcu C ABI (fluid-engine-dev_b200/libjetcu.so), state resident in HBM; `value` = whole-job
  sub-steps/s timed with CUDA events on the library's stream; `e2e` = the same step through the host-buffer
  entry points (set_particles from pinned host memory -> step -> get_particles) with the copies inside the timed
  region; `roofline` = the dominant kernel family (red-black SOR sweeps on the finest MG level) timed live with
  event pairs; `cpu_baseline` = the unmodified reference (oracle/_ref) or the C port on a bounded sample.
  BASELINE configs: default = config "FLIP 256^3" (the metric); --mode apic = config 3; --n 512 --gpus 2|4|8 = config 4;
  scripts/bench_smoke.py = config 5 (its own metric, its own JSON line).
* --impl reference: the reference's own CPU implementation (oracle/_ref/libjetref.so = unmodified Jet, OpenMP, all
  host threads; else the C port) stepping THE SAME scene at THE SAME size: the first k sub-steps that fit the wall budget
  (--ref-budget-s, default 260 s; k <= --steps, no warm-up — every sub-step is minutes of deterministic CPU work).  The line
  carries the ms_per_step it observed, `steps` = k, and per-step PCG iterations / position samples for parity.

A "step" is one GridFluidSolver3::onAdvanceTimeStep sub-step.  Synthetic, deterministic scene.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

def metric_name(args):
    return "%s %d^3 sub-steps/s (MGPCG solve ms and %% of HBM roofline in roofline/stages)" % (args.mode.upper(), args.n)


REF_CACHE = "/tmp/jetcu_reference_arm.json"   # written by --impl reference, read by the own arm's cpu_baseline leg (same box)
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
RELAX_BYTES_PER_CELL_SWEEP = 56.0  # A 32 + b 8 + x 8 r + 8 w, red+black in one residency


def scene(workload, n):
    """Emitter box (lower, upper), initial particle velocity, MG levels for grid n^3 on the unit cube."""
    if workload == "tank":
        # whole footprint filled to y = 15/16 with the reference's h/2 lattice (8 particles/cell; 126.6 M at 256^3 = the
        # "~134 M particles" class), starting at rest.  On particle-built free surfaces at 256^3 the reference's MGPCG (and its
        # replay here) never reaches 1e-9: it leaves through the sigma <= tol^2 test after a handful of iterations with a true
        # residual of ~3e-2 ||b|| (profiles/r02_scene_exploration.md lists 19 scenes — only the box without a
        # free surface converges).  This scene does so with the SAME iteration count every step (7, 5, 5, 5, ... on both the
        # reference and the GPU), so the metric does not swing with the measurement window; the fixed-work keys
        # (ms_per_pcg_iteration, projection.ms_per_iteration) say what an iteration costs regardless.
        return dict(lower=(0.0, 0.0, 0.0), upper=(1.0, 0.9375, 1.0), v0=(0.0, 0.0, 0.0))
    if workload == "slosh":
        # round 1's scene: the same tank with an initial velocity; its first two solves run into the 100-iteration cap
        # (residual 8e11 on the reference and here) and later ones alternate between 3 and 100 iterations
        return dict(lower=(0.0, 0.0, 0.0), upper=(1.0, 0.9375, 1.0), v0=(0.5, 0.0, 0.0))
    if workload == "dam":
        # config 1 scaled: dam column 0.4 x 0.6 x 1.0 (32.3 M particles at 256^3)
        return dict(lower=(0.0, 0.0, 0.0), upper=(0.4, 0.6, 1.0), v0=(0.0, 0.0, 0.0))
    raise SystemExit("unknown workload " + workload)


def mg_levels(n, max_levels=6):
    l = 1
    while l < max_levels and n % 2 == 0 and n > 4:
        n //= 2
        l += 1
    return l


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def read_traffic():
    """dram bytes per launch of the roofline kernel from the committed ncu --set full capture (or null)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def load_oracle(prefer_ref=True):
    from oracle_api import have_port, have_ref, load_port, load_ref
    if prefer_ref and have_ref():
        return load_ref()
    if have_port():
        return load_port()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    return load_port()


MODES = {"pic": 0, "flip": 1, "apic": 2}


def cpu_steps(oracle, workload, n, mode, steps, warmup, dt):
    """Runs the CPU implementation on the scene at n^3; returns (seconds per step list, particles, pcg iterations)."""
    from oracle_api import Grid, HybridParams, MgpcgParams
    sc = scene(workload, n)
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=mg_levels(n)))
    h = oracle.hybrid(prm)
    npart = h.emit_box(sc["lower"], sc["upper"], 0.5 / n, sc["v0"])
    times, iters = [], []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        st = h.step(dt)
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0); iters.append(int(st.pcg_iterations))
    h.close()
    return times, npart, iters


def lattice(lower, upper, spacing):
    """The points VolumeParticleEmitter3 + GridPointGenerator3 emit for a box inside the domain, in their order (k, j, i;
    src/jet/grid_point_generator3.cpp:12-37: position = idx * spacing + lower while idx * spacing <= extent) — the same set
    jetcu_hybrid_emit_box / the reference's emitter produce (tests/test_gpu_particles.py::test_emit_box_bit_exact), built with
    numpy because the reference's emitter alone spends two minutes hashing 126 M points for its overlap test."""
    import numpy as np
    def axis(lo, hi):
        m = int(np.floor((hi - lo) / spacing)) + 2
        i = np.arange(m, dtype=np.float64)
        return (i * spacing + lo)[i * spacing <= (hi - lo)]
    xs, ys, zs = (axis(lower[a], upper[a]) for a in range(3))
    out = np.empty((len(zs), len(ys), len(xs), 3))
    out[..., 0] = xs[None, None, :]; out[..., 1] = ys[None, :, None]; out[..., 2] = zs[:, None, None]
    return out.reshape(-1, 3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU reference; the others exit 0 without work
    import numpy as np
    from oracle_api import Grid, HybridParams, MgpcgParams
    oracle = load_oracle(True)
    oracle.set_num_threads(0)
    cores = oracle.num_threads()
    n = args.n
    sc = scene(args.workload, n)
    t_start = time.perf_counter()
    prm = HybridParams.make(Grid.cube(n), mode=MODES[args.mode], mgpcg=MgpcgParams.make(max_iterations=100, max_levels=mg_levels(n)))
    h = oracle.hybrid(prm)
    if args.workload in ("tank", "slosh"):   # box faces on representable lattice planes: every lattice point is emitted
        pts = lattice(sc["lower"], sc["upper"], 0.5 / n)
        npart = len(pts)
        h.set_particles(pts, np.broadcast_to(np.asarray(sc["v0"], dtype=np.float64), pts.shape).copy())
        del pts
    else:                                    # the emitter's own on-face rounding decides (tests: test_emit_box_face_rounding)
        npart = h.emit_box(sc["lower"], sc["upper"], 0.5 / n, sc["v0"])
    setup_s = time.perf_counter() - t_start
    idx = np.arange(0, npart, 65521)
    times, iters, resid, samples = [], [], [], []
    # the first k sub-steps that fit the wall budget; never fewer than one, never more than --steps
    while len(times) < max(args.steps, 1):
        if times and (time.perf_counter() - t_start) + 1.15 * max(times) > args.ref_budget_s:
            break
        t0 = time.perf_counter()
        st = h.step(args.dt)
        times.append(time.perf_counter() - t0)
        iters.append(int(st.pcg_iterations)); resid.append(float(st.pcg_last_residual))
        pos, _ = h.get_particles()
        samples.append([float(x) for x in pos[idx[:4]].ravel()] + [float(pos[:, a].sum()) for a in range(3)])
        del pos
    h.close()
    k = len(times)
    sec = sum(times) / k
    value = 1.0 / sec
    sample = ("the %s scene itself at %d^3 (%d particles), the first %d sub-steps from "
              "rest, no warm-up, %.1f s each on %d threads (wall budget %.0f s, setup %.1f s)" %
              (args.workload, n, npart, k, sec, cores, args.ref_budget_s, setup_s))
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": k, "steps_requested": args.steps, "warmup": 0, "warmup_requested": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, npart),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": oracle.kind, "sample": sample,
                         "seconds_per_step": times, "pcg_iterations": iters, "pcg_last_residual": resid,
                         "wall_s": time.perf_counter() - t_start},
        "parity_trace": {"particles_0_3_xyz_then_position_sums": samples},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        json.dump({"config": line["config"], "cpu_baseline": line["cpu_baseline"], "when": time.time()}, open(REF_CACHE, "w"))
    except OSError:
        pass
    print(json.dumps(line), flush=True)


def workload_config(args, npart):
    sc = scene(args.workload, args.n)
    names = {"tank": "tank at rest", "slosh": "sloshing tank", "dam": "dam break"}
    c = {"workload": "%s %d^3 %s, h/2 lattice fill %s-%s, v0 %s, dt %.6g s, MGPCG(100, %d levels, 5/5/20/20, 1e-9, "
                     "omega 1.5, red-black)" % (args.mode.upper(), args.n, names[args.workload], sc["lower"], sc["upper"], sc["v0"],
                                                args.dt, mg_levels(args.n)),
         "grid": [args.n] * 3, "cache": "inputs larger than L2 (particle + grid state >> 126 MB)",
         "parallelism": "1 GPU" if args.gpus == 1 else
         "%d GPUs, z-slabs of %d planes per rank (one domain, strong scaling): halo-plane exchange per smoother block / matvec / "
         "transfer, slab-invariant PCG dot products, coarse MG levels replicated, particle migration + 2 ghost planes per step" %
         (args.gpus, args.n // args.gpus)}
    if npart is not None:
        c["particles"] = int(npart)
    return c


def run_projection(jc, n, peak):
    """BASELINE config 2: standalone GridFractionalSinglePhasePressureSolver3 + MGPCG projection at n^3 on the flat-slab
    input with the box collider (the variant on which the reference's own solver converges), through
    jetcu_pressure_solve with host arrays.  solve_ms = device time of the MGPCG solve (CUDA events inside the library);
    bytes = SURVEY §8d's model: N * [(56 + 1647 + 16) + iterations * 1799]."""
    import scenes
    from fluid_engine_dev_b200.jetcu import MgpcgParams
    sc = scenes.projection_scene(n, "slab", True)
    levels = mg_levels(n, 7)
    prm = MgpcgParams.make(max_iterations=100, max_levels=levels)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        _, _, _, _, res = jc.pressure_solve(sc["grid"], sc["u"], sc["v"], sc["w"], prm, collider_sdf=sc["collider_sdf"],
                                            fluid_sdf=sc["fluid_sdf"], want_pressure=True)
        call_ms = 1e3 * (time.perf_counter() - t0)
        if best is None or res.solve_ms < best[0].solve_ms:
            best = (res, call_ms)
    res, call_ms = best
    ncell = float(n) ** 3
    model_bytes = ncell * ((56 + 1647 + 16) + res.iterations * 1799.0)
    gbs = model_bytes / (res.solve_ms * 1e-3) / 1e9 if res.solve_ms > 0 else None
    return {"workload": "config 2: projection %d^3, fluidSdf = y - 0.5, v = +1, box collider, MGPCG(100, %d levels, "
                        "5/5/20/20, 1e-9, omega 1.5, red-black)" % (n, levels),
            "mgpcg_solve_ms": res.solve_ms, "iterations": int(res.iterations),
            "ms_per_iteration": res.solve_ms / max(int(res.iterations), 1),
            "last_residual": res.last_residual, "true_rel_residual": res.true_residual_l2 / max(res.b_l2, 1e-300),
            "call_ms_host_buffers": call_ms, "model_bytes": model_bytes, "achieved_gbs": gbs,
            "frac_of_hbm_peak": (gbs / peak) if gbs else None}


def run_jetcu(args):
    import numpy as np
    import torch
    from jetcu_loader import load_package
    load_package()
    from fluid_engine_dev_b200 import Jetcu
    from fluid_engine_dev_b200.jetcu import Grid, HybridParams, MgpcgParams
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    jc = Jetcu(local)
    if world > 1:
        jc.comm_init_torch()   # z-slab communicator (NCCL) next to torch's own process group
    stream = torch.cuda.ExternalStream(jc.stream(), device=dev)
    n = args.n
    sc = scene(args.workload, n)
    levels = mg_levels(n)
    mode = MODES[args.mode]
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=levels))
    dt = args.dt
    parity = slab_parity_check(jc, dist, dev, rank, world, local, mode) if world > 1 else None

    def fresh_scene():
        """Emits the scene and runs the W untimed warm-up steps: both timed passes start from the same state."""
        hh = jc.hybrid(prm)
        count = hh.emit_box(sc["lower"], sc["upper"], 0.5 / n, sc["v0"])
        for _ in range(args.warmup):
            hh.step(dt)
        jc.synchronize()
        return hh, count

    h, npart = fresh_scene()

    def prof(kind):
        ms, la, un = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
        jc._call("profile_get", C.c_int(kind), C.byref(ms), C.byref(la), C.byref(un))
        return ms.value, la.value, un.value

    # ---------------------------------------------------------------- timed region: device-resident steps
    clocks = ClockSampler(local) if rank == 0 else None
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    l0 = jc.launch_count()
    m0, b0 = jc.comm_stats()
    t0 = time.perf_counter()
    e0.record(stream)
    stats = []
    for _ in range(args.steps):
        stats.append(h.step(dt).as_dict())
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    wall = time.perf_counter() - t0
    launches = jc.launch_count() - l0
    m1, b1 = jc.comm_stats()
    dev_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    clk = clocks.stop() if clocks else None
    steps_per_s = args.steps / dev_s   # one domain split over the ranks: the job advances `steps` sub-steps

    # ---------------------------------------------------------------- e2e: host buffers in and out every step
    from fluid_engine_dev_b200.jetcu import _p, u64, PU64
    k_e2e = max(0, min(args.steps, args.e2e_steps))
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    bytes_in = bytes_out = 0
    e2e_note = "skipped (--e2e-steps 0)"
    t0 = time.perf_counter()
    if k_e2e == 0:
        pass
    elif world == 1:
        pos, vel = h.get_particles()
        nloc = int(npart)
        ppos = pin(pos.shape, torch.float64); ppos[...] = pos
        pvel = pin(vel.shape, torch.float64); pvel[...] = vel
        opos, ovel = pin(pos.shape, torch.float64), pin(vel.shape, torch.float64)
        del pos, vel
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            jc._call_raw("hybrid_set_particles", h.h, u64(npart), _p(ppos), _p(pvel))
            h.step(dt)
            jc._call_raw("hybrid_get_particles", h.h, _p(opos), _p(ovel))
            ppos, opos = opos, ppos
            pvel, ovel = ovel, pvel
        torch.cuda.synchronize(); barrier()
        bytes_in = bytes_out = nloc * 48
        e2e_note = ("jetcu_hybrid_set_particles(pinned host) -> jetcu_hybrid_step -> jetcu_hybrid_get_particles(host), "
                    "wall clock with synchronize on both sides")
    else:
        # every rank round-trips ITS OWN particles (ids + positions + velocities) through pinned host memory
        capn = int(npart // world * 1.25) + 1024
        ids = pin((capn,), torch.int64).view(np.uint64)
        ppos, pvel = pin((capn, 3), torch.float64), pin((capn, 3), torch.float64)
        nloc = h.num_local_particles()
        assert nloc <= capn
        jc._call_raw("hybrid_get_local_particles", h.h, _p(ids, PU64), _p(ppos), _p(pvel))
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        moved = 0
        for _ in range(k_e2e):
            jc._call_raw("hybrid_set_local_particles", h.h, u64(nloc), _p(ids, PU64), _p(ppos), _p(pvel), u64(npart))
            h.step(dt)
            moved += nloc
            nloc = h.num_local_particles()
            assert nloc <= capn
            jc._call_raw("hybrid_get_local_particles", h.h, _p(ids, PU64), _p(ppos), _p(pvel))
            moved += nloc
        torch.cuda.synchronize(); barrier()
        bytes_in = bytes_out = int(moved // (2 * k_e2e)) * 56
        e2e_note = ("per rank: jetcu_hybrid_set_local_particles(pinned host) -> jetcu_hybrid_step -> "
                    "jetcu_hybrid_get_local_particles(host); bytes are rank 0's, wall clock max over ranks")
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_rate = (k_e2e / e2e_s) if k_e2e else None
    h.close()

    # ---------------------------------------------------------------- roofline pass: the same K steps again with the
    # library's per-kernel-family CUDA-event pairs switched on.  The pairs cannot live inside the CUDA graph the V-cycle
    # normally replays from, so this pass launches the same kernels eagerly; share_of_step is relative to THIS pass.
    h, _ = fresh_scene()   # same particles, same warm-up: the profiled steps are the timed steps over again
    jc._call("profile_enable", C.c_int(1))
    p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    p0.record(stream)
    prof_stats = []
    for _ in range(args.steps):
        prof_stats.append(h.step(dt).pcg_iterations)
    p1.record(stream)
    torch.cuda.synchronize(); barrier()
    prof_s = p0.elapsed_time(p1) * 1e-3
    relax_ms, relax_launches, relax_units = prof(0)
    mvm_ms, mvm_launches, mvm_units = prof(2)
    jc._call("profile_enable", C.c_int(0))
    prof_iters = [int(x) for x in prof_stats]
    h.close()

    if rank != 0:
        teardown(jc, dist)
        return

    peak, peak_src = read_peaks()
    ncell = n ** 3
    relax_gbs = (relax_units * RELAX_BYTES_PER_CELL_SWEEP / (relax_ms * 1e-3) / 1e9) if relax_ms > 0 else None
    traffic = read_traffic()
    roof = {"bound": "hbm", "kernel": "red-black SOR sweep, finest MG level (k_relax_*)",
            "achieved": relax_gbs, "peak": peak, "unit": "GB/s", "frac": (relax_gbs / peak) if relax_gbs else None,
            "peak_source": peak_src,
            # DRAM bytes per launch from the committed ncu capture: bytes per cell and colour pass x the cells one launch covers
            "traffic": ((traffic or {}).get("relax_l0_dram_bytes_per_cell_pass", 0.0) * relax_units * 2.0 / relax_launches)
            if (traffic and relax_launches) else None,
            "traffic_source": (traffic or {}).get("source"),
            "bytes_per_launch_algorithmic": (relax_units * RELAX_BYTES_PER_CELL_SWEEP / relax_launches) if relax_launches else None,
            "launches": relax_launches, "avg_launch_ms": (relax_ms / relax_launches) if relax_launches else None,
            "share_of_step": relax_ms * 1e-3 / prof_s if prof_s > 0 else None,
            "timed_in": "second pass over the same scene, warm-up and %d steps with per-family CUDA-event pairs (eager launches; "
                        "the value pass replays the V-cycle as a CUDA graph): %.3f ms/step, PCG iterations %s" %
                        (args.steps, 1e3 * prof_s / args.steps, prof_iters),
            "mvm_l0_gbs": (mvm_units * 48.0 / (mvm_ms * 1e-3) / 1e9) if mvm_ms > 0 else None}
    if relax_gbs and traffic and traffic.get("relax_l0_dram_bytes_per_cell_pass"):
        # `achieved` counts SURVEY 8d's ALGORITHMIC bytes (28 B per cell and colour pass).  The skewed z-chunk wavefront re-sweeps
        # a plane while it is still in the L2, so part of those bytes never reach DRAM and the algorithmic rate may exceed the
        # DRAM copy peak (frac > 1); the DRAM-side rate follows from the ncu capture's bytes per cell and pass.
        dram_b = float(traffic["relax_l0_dram_bytes_per_cell_pass"])
        roof["dram_side_gbs_estimate"] = relax_gbs * dram_b / (RELAX_BYTES_PER_CELL_SWEEP / 2.0)
        roof["dram_side_frac_estimate"] = roof["dram_side_gbs_estimate"] / peak
        roof["note"] = ("achieved = algorithmic bytes (28 B per cell and colour pass) / measured time; the L2-resident wavefront serves "
                        "part of them from the L2 (ncu: %.1f B of DRAM traffic per cell and pass), which is how frac can exceed 1" % dram_b)
    mean = lambda k: sum(s[k] for s in stats) / len(stats)
    stages = {k: mean(k) for k in ("ms_total", "ms_bin", "ms_p2g", "ms_sdf", "ms_extrapolate", "ms_bc", "ms_pressure",
                                   "ms_g2p", "ms_move", "ms_mgpcg_solve")}
    stages["pcg_iterations"] = [int(s["pcg_iterations"]) for s in stats]
    stages["pcg_true_rel_residual"] = [s["pcg_true_residual_l2"] / max(s["pcg_b_l2"], 1e-300) for s in stats]
    # per-stage achieved bandwidth on SURVEY §8d's algorithmic bytes (P = particles, N = cells; FLIP): what the stage must
    # move at the very least, divided by its measured time — the distance of every stage from the HBM roofline, not only
    # the dominant kernel's.  At N > 1 the times are rank 0's and the bytes the whole domain's divided by the ranks.
    P_, N_ = float(npart) / world, float(ncell) / world
    mean_iters = sum(stages["pcg_iterations"]) / max(len(stages["pcg_iterations"]), 1)
    # SURVEY §8d per particle: reorder gather 48 r + 48 w (APIC 120 + 120), P2G 48 (APIC 120), G2P PIC 24 r + 24 w / FLIP 48 r +
    # 24 w / APIC 24 r + 96 w, move 48 r + 48 w; per cell: P2G 51 grid writes, G2P 24 (+ 12 f32 FLIP), move 24
    apic = args.mode == "apic"
    attr = 120.0 if apic else 48.0
    g2p_p = {"pic": 48.0, "flip": 72.0, "apic": 120.0}[args.mode]
    g2p_c = 36.0 if args.mode == "flip" else 24.0
    stage_bytes = {"ms_bin": (40.0 + 2.0 * attr) * P_, "ms_p2g": attr * P_ + 51.0 * N_, "ms_sdf": 24.0 * P_ + 8.0 * N_,
                   "ms_g2p": g2p_p * P_ + g2p_c * N_, "ms_move": 96.0 * P_ + 24.0 * N_,
                   "ms_pressure": N_ * ((56.0 + 1647.0 + 16.0) + mean_iters * 1799.0)}
    # PIC / FLIP: G2P and advection are one kernel (hybrid.cu: hyb_g2p_move) — ms_g2p is then the gap between two event records
    fused_adv = stages.get("ms_g2p", 0) < 0.05
    if fused_adv:   # one pass over the particles: G2P's velocity write and the advection's re-read disappear
        stage_bytes["ms_move"] = (48.0 + 48.0) * P_ + (g2p_c + 24.0) * N_
        del stage_bytes["ms_g2p"]
    stage_roofline = {}
    for k, bts in stage_bytes.items():
        if stages.get(k, 0) > 0:
            gbs = bts / (stages[k] * 1e-3) / 1e9
            stage_roofline[k[3:]] = {"algorithmic_gb": bts / 1e9, "gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    if "move" in stage_roofline:
        # ncu (profiles/r02_move_variants.md): the advection kernel follows its FP64 instruction count — 10 trilinear FP64 samples
        # of 3 components per particle replayed operation by operation (no FMA contraction: bit-identical positions) are
        # ~1 100 FP64 instructions per particle; neither occupancy nor ten times fewer loads (corner cache) move it.
        # floor = P * 1100 / (148 SMs * 64 FP64 lanes * clock)
        fp64_floor_ms = P_ * 1100.0 / (148 * 64 * 1.9e9) * 1e3
        stage_roofline["move"].update({"bound": "fp64 issue (not HBM)", "fp64_floor_ms": fp64_floor_ms,
                                       "frac_of_fp64_floor": fp64_floor_ms / stages["ms_move"],
                                       "fused_with_g2p": bool(fused_adv)})
    # fixed-work numbers (independent of how many iterations a step's solve happens to take): one PCG iteration = one
    # V-cycle (5/5/20/20 sweeps) + matvec + 2 dots + 3 axpy; the solve's prologue is one more V-cycle
    tot_it = sum(int(s["pcg_iterations"]) + 1 for s in stats)
    ms_per_it = sum(s["ms_mgpcg_solve"] for s in stats) / max(tot_it, 1)
    fixed = {"ms_per_pcg_iteration": ms_per_it, "iterations_plus_prologues": tot_it,
             "pcg_iteration_gbs": (N_ * 1799.0 / (ms_per_it * 1e-3) / 1e9) if ms_per_it > 0 else None,
             "ms_particle_stages": sum(stages[k] for k in ("ms_bin", "ms_p2g", "ms_sdf", "ms_g2p", "ms_move")),
             "ms_grid_stages": stages["ms_extrapolate"] + stages["ms_bc"] + stages["ms_pressure"] - stages["ms_mgpcg_solve"]}
    if fixed["pcg_iteration_gbs"]:
        fixed["pcg_iteration_frac_of_hbm_peak"] = fixed["pcg_iteration_gbs"] / peak
    line = {
        "metric": metric_name(args), "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, npart),
        "mcell_per_s": steps_per_s * ncell / 1e6, "particles_per_s": steps_per_s * npart,
        "mgpcg_solve_ms": stages["ms_mgpcg_solve"], "stages_ms": stages, "stage_roofline": stage_roofline, "fixed_work": fixed,
        "wall_s_timed_region": wall,
        "gpu_launches": int(launches), "clocks": clk,
        "e2e": {"value": e2e_rate, "unit": "steps/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
                "steps": k_e2e, "note": e2e_note},
        "roofline": roof,
    }
    if parity is not None:
        line["parity_check"] = parity
    if world > 1:
        line["comm"] = {"rank0_messages_per_step": (m1 - m0) / args.steps, "rank0_bytes_sent_per_step": (b1 - b0) / args.steps,
                        "backend": ("peer-memory mailbox kernels over NVLink (CUDA IPC) for halo planes, NCCL all-gather"
                                    if jc.comm_p2p_enabled() else "NCCL send/recv + all-gather on the solver stream")}
    # the two legs below are reported next to the measurement, they are not part of it: a failure in one of them is recorded
    # in its own key and must not cost the line
    if world == 1 and not args.no_projection:
        try:
            line["projection"] = run_projection(jc, args.n, peak)
        except Exception as e:  # noqa: BLE001
            line["projection"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_leg(args, n, dt)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line), flush=True)
    teardown(jc, dist)


def cpu_baseline_leg(args, n, dt):
    """The reference's CPU path on the box's host cores (reported, not the target).  If `--impl reference` ran on this box
    for the same config (the driver runs it right before this arm), its measurement of the REAL workload is reported;
    otherwise a bounded sample: the same scene at 96^3, scaled to the metric's unit by the cell ratio."""
    try:
        c = json.load(open(REF_CACHE))
        if c["config"] == workload_config(args, c["config"].get("particles")) and time.time() - c["when"] < 6 * 3600:
            out = dict(c["cpu_baseline"])
            out["sample"] = "measured by `bench.py --impl reference` on this box %.0f min ago: %s" % (
                (time.time() - c["when"]) / 60.0, out["sample"])
            return out
    except Exception:  # noqa: BLE001
        pass
    oracle = load_oracle(True)
    oracle.set_num_threads(0)
    n_s = args.sample_n or 96
    times, np_s, iters = cpu_steps(oracle, args.workload, n_s, MODES[args.mode], 2, 1, dt * (n / n_s))
    ratio = (n_s / float(n)) ** 3
    sec = sum(times) / len(times)
    return {
        "value": ratio / sec, "unit": "steps/s", "cores": oracle.num_threads(), "kind": oracle.kind,
        "sample": "%s scene at %d^3 (%d particles): 1 warm-up + %d timed sub-steps of %.2f s each (%.1f s of CPU work); "
                  "value = sample steps/s x (%d/%d)^3 cell ratio (a bounded sample: `--impl reference` times the %d^3 scene "
                  "itself)" % (args.workload, n_s, np_s, len(times), sec, sum(times), n_s, n, n),
        "sample_pcg_iterations": iters}


def slab_parity_check(jc, dist, dev, rank, world, local, mode, n=64, steps=3):
    """Evidence that the z-slab decomposition is bit-identical to the undivided domain, taken inside the measured job (the
    driver's GPU-test box has one GPU): the dam break at 64^3 is stepped on the `world` slabs and, on rank 0, once more on a
    single-GPU context; positions, velocities and pressure must agree bit for bit."""
    import numpy as np
    import torch
    from fluid_engine_dev_b200 import Jetcu
    from fluid_engine_dev_b200.jetcu import Grid, HybridParams, MgpcgParams
    if n % world or (n // world) % 2:
        return {"skipped": "%d^3 does not split into %d even slabs" % (n, world)}
    prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=mg_levels(n)))
    h = jc.hybrid(prm)
    npart = h.emit_box((0.0, 0.0, 0.0), (0.4, 0.6, 1.0), 0.5 / n)
    its = [int(h.step(1.0 / 60.0).pcg_iterations) for _ in range(steps)]
    pos, vel = h.get_particles()     # this rank's particles, zeros elsewhere: the sum over ranks is the whole set
    pr = h.get_pressure()
    h.close()
    t = torch.from_numpy(np.concatenate([pos.ravel(), vel.ravel()])).to(dev)
    dist.all_reduce(t)
    out = None
    if rank == 0:
        whole = t.cpu().numpy()
        solo = Jetcu(local)          # a second context without a communicator: the undivided domain
        hs = solo.hybrid(prm)
        assert hs.emit_box((0.0, 0.0, 0.0), (0.4, 0.6, 1.0), 0.5 / n) == npart
        its1 = [int(hs.step(1.0 / 60.0).pcg_iterations) for _ in range(steps)]
        p1, v1 = hs.get_particles()
        pr1 = hs.get_pressure()
        hs.close(); solo.close()
        ref = np.concatenate([p1.ravel(), v1.ravel()])
        out = {"scene": "FLIP-mode-%d dam break %d^3, %d particles, %d sub-steps: %d z-slabs vs one GPU" % (mode, n, npart, steps, world),
               "bit_identical": bool(np.array_equal(whole, ref) and np.array_equal(pr, pr1) and its == its1),
               "max_abs_diff_particles": float(np.abs(whole - ref).max()), "max_abs_diff_pressure": float(np.abs(pr - pr1).max()),
               "pcg_iterations": its, "pcg_iterations_one_gpu": its1}
    dist.barrier()
    return out


def teardown(jc, dist):
    """Release the solver (graphs first, then the NCCL communicator) and the process group.  Every measurement is done and
    rank 0's line is out by now, so a communicator that refuses to die must not cost the run: a watchdog ends the process
    with status 0 if the orderly teardown has not finished in 60 s (ncclCommDestroy waits for peers and for captured
    graphs; it has no timeout of its own)."""
    import threading
    t = threading.Timer(60.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    jc.close()
    if dist is not None:
        dist.destroy_process_group()
    t.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="jetcu", choices=["jetcu", "reference"])
    ap.add_argument("--workload", default="tank", choices=["tank", "slosh", "dam"])
    ap.add_argument("--mode", default="flip", choices=["flip", "apic", "pic"])
    ap.add_argument("--ref-budget-s", type=float, default=float(os.environ.get("JETCU_REF_BUDGET_S", "260")),
                    help="--impl reference: wall budget for the real-size CPU run")
    ap.add_argument("--n", "--grid", dest="n", type=int, default=256,
                    help="grid size n^3 (use --grid under torchrun: its own parser finds --n ambiguous)")
    ap.add_argument("--dt", type=float, default=1.0 / 240.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sample-n", type=int, default=0, help="grid size of the CPU sample (0 = automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-projection", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_jetcu(args)


if __name__ == "__main__":
    main()
