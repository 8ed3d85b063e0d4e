#!/usr/bin/env python
"""bench.py — FLIP 256^3 sub-steps/s of the B200-native hybrid-liquid path (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl jetcu|reference] [--workload tank|dam] [--n 256]

* own arm (default): jetcu C ABI (fluid-engine-dev_b200/libjetcu.so), state resident in HBM; `value` = whole-job
  sub-steps/s timed with CUDA events on the library's stream; `e2e` = the same step through the host-buffer
  entry points (set_particles from pinned host memory -> step -> get_particles) with the copies inside the timed
  region; `roofline` = the dominant kernel family (red-black SOR sweeps on the finest MG level) timed live with
  event pairs inside the same timed region; `cpu_baseline` = the unmodified reference (oracle/_ref) or the C port
  on a bounded sample of the same scene.
* --impl reference: the reference's own CPU implementation (oracle/_ref/libjetref.so = unmodified Jet, OpenMP, all
  host threads; else the C port) stepping the same scene at a reduced grid (bounded sample), converted to the
  256^3-equivalent rate by the cell ratio.

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

METRIC = "FLIP 256^3 sub-steps/s (MGPCG solve ms and % of HBM roofline in roofline/stages)"
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
RELAX_BYTES_PER_CELL_SWEEP = 56.0  # A 32 + b 8 + x 8 r + 8 w, red+black in one residency


def scene(workload, n):
    """Emitter box (lower, upper), initial particle velocity, MG levels for grid n^3 on the unit cube."""
    if workload == "tank":
        # whole footprint filled to y = 15/16 with the reference's h/2 lattice (8 particles/cell; 126.6 M at 256^3 =
        # the "~134 M particles" class), sloshing start.  Flat free surface: the reference MGPCG converges on it
        # (the 256^3 dam column makes the reference's own solver exit early at residual ~1).
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


def pick_sample_n(total_steps, budget_s=200.0):
    # measured on 8-16 host threads: a tank step costs ~1.6e-5 s per cell (pressure dominated, mostly serial)
    for n in (96, 64, 48, 32):
        if total_steps * 1.6e-5 * n ** 3 <= budget_s:
            return n
    return 32


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU reference; the others exit 0 without work
    from oracle_api import FLIP
    oracle = load_oracle(True)
    oracle.set_num_threads(0)
    cores = oracle.num_threads()
    n_s = args.sample_n or pick_sample_n(args.steps + args.warmup)
    dt = args.dt * (args.n / n_s)  # same CFL number as the full-size run
    t0 = time.perf_counter()
    times, npart, iters = cpu_steps(oracle, args.workload, n_s, FLIP, args.steps, args.warmup, dt)
    wall = time.perf_counter() - t0
    sec = sum(times) / len(times)
    ratio = (n_s / float(args.n)) ** 3
    value = ratio / sec
    sample = ("%s scene at %d^3 (%d particles), %d timed sub-steps of %.2f s each on %d threads; value = sample "
              "steps/s x (%d/%d)^3 cell ratio (linear-cost conversion to the %d^3 workload)" %
              (args.workload, n_s, npart, len(times), sec, cores, n_s, args.n, args.n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / ratio, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, None),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": oracle.kind, "sample": sample,
                         "sample_seconds_per_step": sec, "sample_pcg_iterations": iters, "wall_s": wall},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, npart):
    sc = scene(args.workload, args.n)
    c = {"workload": "FLIP %d^3 %s, h/2 lattice fill %s-%s, v0 %s, dt %.6g s, MGPCG(100, %d levels, 5/5/20/20, 1e-9, "
                     "omega 1.5, red-black)" % (args.n, "sloshing tank" if args.workload == "tank" else "dam break",
                                                sc["lower"], sc["upper"], sc["v0"], args.dt, mg_levels(args.n)),
         "grid": [args.n] * 3, "cache": "inputs larger than L2 (particle + grid state >> 126 MB)",
         "parallelism": "1 GPU" if args.gpus == 1 else
         "%d GPUs, z-slabs of %d planes per rank (one domain, strong scaling): NCCL halo-plane exchange per smoother colour "
         "pass / matvec / transfer, slab-invariant PCG dot products, coarse MG levels replicated, particle migration + "
         "2 ghost planes per step" % (args.gpus, args.n // args.gpus)}
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
    from fluid_engine_dev_b200.jetcu import FLIP, Grid, HybridParams, MgpcgParams
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
    prm = HybridParams.make(Grid.cube(n), mode=FLIP, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=levels))
    dt = args.dt

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
    k_e2e = max(1, min(args.steps, args.e2e_steps))
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    if world == 1:
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
        capn = int(npart // world * 1.5) + 1024
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
    e2e_rate = k_e2e / e2e_s
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
            "peak_source": peak_src, "traffic": (traffic or {}).get("relax_l0_bytes_per_launch"),
            "traffic_source": (traffic or {}).get("source"),
            "bytes_per_launch_algorithmic": (relax_units * RELAX_BYTES_PER_CELL_SWEEP / relax_launches) if relax_launches else None,
            "launches": relax_launches, "avg_launch_ms": (relax_ms / relax_launches) if relax_launches else None,
            "share_of_step": relax_ms * 1e-3 / prof_s if prof_s > 0 else None,
            "timed_in": "second pass over the same scene, warm-up and %d steps with per-family CUDA-event pairs (eager launches; "
                        "the value pass replays the V-cycle as a CUDA graph): %.3f ms/step, PCG iterations %s" %
                        (args.steps, 1e3 * prof_s / args.steps, prof_iters),
            "mvm_l0_gbs": (mvm_units * 48.0 / (mvm_ms * 1e-3) / 1e9) if mvm_ms > 0 else None}
    mean = lambda k: sum(s[k] for s in stats) / len(stats)
    stages = {k: mean(k) for k in ("ms_total", "ms_bin", "ms_p2g", "ms_sdf", "ms_extrapolate", "ms_bc", "ms_pressure",
                                   "ms_g2p", "ms_move")}
    stages["pcg_iterations"] = [int(s["pcg_iterations"]) for s in stats]
    stages["pcg_true_rel_residual"] = [s["pcg_true_residual_l2"] / max(s["pcg_b_l2"], 1e-300) for s in stats]
    # per-stage achieved bandwidth on SURVEY §8d's algorithmic bytes (P = particles, N = cells; FLIP): what the stage must
    # move at the very least, divided by its measured time — the distance of every stage from the HBM roofline, not only
    # the dominant kernel's.  At N > 1 the times are rank 0's and the bytes the whole domain's divided by the ranks.
    P_, N_ = float(npart) / world, float(ncell) / world
    mean_iters = sum(stages["pcg_iterations"]) / max(len(stages["pcg_iterations"]), 1)
    stage_bytes = {"ms_bin": (40.0 + 96.0) * P_, "ms_p2g": 48.0 * P_ + 51.0 * N_, "ms_sdf": 24.0 * P_ + 8.0 * N_,
                   "ms_g2p": 72.0 * P_ + 36.0 * N_, "ms_move": 96.0 * P_ + 24.0 * N_,
                   "ms_pressure": N_ * ((56.0 + 1647.0 + 16.0) + mean_iters * 1799.0)}
    stage_roofline = {}
    for k, bts in stage_bytes.items():
        if stages.get(k, 0) > 0:
            gbs = bts / (stages[k] * 1e-3) / 1e9
            stage_roofline[k[3:]] = {"algorithmic_gb": bts / 1e9, "gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    line = {
        "metric": METRIC, "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, npart),
        "mcell_per_s": steps_per_s * ncell / 1e6, "particles_per_s": steps_per_s * npart,
        "mgpcg_solve_ms": stages["ms_pressure"], "stages_ms": stages, "stage_roofline": stage_roofline,
        "wall_s_timed_region": wall,
        "gpu_launches": int(launches), "clocks": clk,
        "e2e": {"value": e2e_rate, "unit": "steps/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
                "steps": k_e2e, "note": e2e_note},
        "roofline": roof,
    }
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
    """The reference's CPU path on the box's host cores, on a bounded sample of the same workload (reported, not the target)."""
    from oracle_api import FLIP
    oracle = load_oracle(True)
    oracle.set_num_threads(0)
    n_s = args.sample_n or 64
    times, np_s, iters = cpu_steps(oracle, args.workload, n_s, FLIP, 3, 1, dt * (n / n_s))
    ratio = (n_s / float(n)) ** 3
    sec = sum(times) / len(times)
    return {
        "value": ratio / sec, "unit": "steps/s", "cores": oracle.num_threads(), "kind": oracle.kind,
        "sample": "%s scene at %d^3 (%d particles): 1 warm-up + %d timed sub-steps of %.2f s each (%.1f s of CPU work); "
                  "value = sample steps/s x (%d/%d)^3 cell ratio" % (args.workload, n_s, np_s, len(times), sec,
                                                                     sum(times), n_s, n),
        "sample_pcg_iterations": iters}


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
    ap.add_argument("--workload", default="tank", choices=["tank", "dam"])
    ap.add_argument("--n", type=int, default=256)
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
