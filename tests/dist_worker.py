"""Worker of the multi-GPU parity test (launched by tests/test_gpu_slabs.py with torchrun, one rank per GPU).

Every rank runs the SAME workload twice: z-slab distributed over the communicator, and alone on its own GPU through a
second, communicator-less context.  The distributed result must equal the single-GPU result BIT FOR BIT
(include/jetcu.h: slab-invariant dot products, integer P2G accumulators, exact min for the particle SDF).
Prints one JSON line per rank; exit code 0 iff every check passed on this rank."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from jetcu_loader import load_package
    load_package()
    from fluid_engine_dev_b200 import Jetcu
    from fluid_engine_dev_b200.jetcu import APIC, FLIP, Grid, HybridParams, MgpcgParams
    import scenes

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    out = {"rank": rank, "world": world, "n": n, "checks": {}}
    ok = True

    def check(name, cond, info=None):
        nonlocal ok
        out["checks"][name] = bool(cond) if info is None else [bool(cond), info]
        ok = ok and bool(cond)

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dist.all_reduce(t)
        return t.cpu().numpy()

    jd = Jetcu(local)          # distributed context
    jd.comm_init_torch()
    js = Jetcu(local)          # single-GPU context on the same device: the bit-exact yardstick
    check("comm", jd.lib.jetcu_comm_world(jd.ctx) == world and jd.lib.jetcu_comm_rank(jd.ctx) == rank)

    # ---- seam 2: projection (slab scene + box collider), distributed vs alone
    sc = scenes.projection_scene(n, "slab", True)
    prm = MgpcgParams.make(max_iterations=100, max_levels=6)
    ud, vd, wd, pd, rd = jd.pressure_solve(sc["grid"], sc["u"], sc["v"], sc["w"], prm, collider_sdf=sc["collider_sdf"],
                                           fluid_sdf=sc["fluid_sdf"])
    us, vs, ws, ps, rs = js.pressure_solve(sc["grid"], sc["u"], sc["v"], sc["w"], prm, collider_sdf=sc["collider_sdf"],
                                           fluid_sdf=sc["fluid_sdf"])
    check("projection_iterations", rd.iterations == rs.iterations, [int(rd.iterations), int(rs.iterations)])
    check("projection_last_residual", rd.last_residual == rs.last_residual, [rd.last_residual, rs.last_residual])
    check("projection_pressure_bits", np.array_equal(pd, ps), float(np.abs(pd - ps).max()))
    check("projection_velocity_bits", np.array_equal(ud, us) and np.array_equal(vd, vs) and np.array_equal(wd, ws))
    check("projection_converged", rs.true_residual_l2 <= 1e-8 * rs.b_l2, rs.true_residual_l2 / max(rs.b_l2, 1e-300))
    out["projection_ms"] = {"slabs": rd.solve_ms, "alone": rs.solve_ms}

    # ---- seam 1: MGPCG on a host-supplied Poisson hierarchy
    levels = 4
    A = scenes.poisson_levels(n, levels)
    rng = np.random.default_rng(7)
    b = rng.uniform(-1, 1, (n, n, n)); b -= b.mean()
    xd, r1 = jd.mgpcg_solve(A, b, MgpcgParams.make(max_iterations=30, max_levels=levels))
    xs, r2 = js.mgpcg_solve(A, b, MgpcgParams.make(max_iterations=30, max_levels=levels))
    check("mgpcg_bits", np.array_equal(xd, xs) and r1.iterations == r2.iterations, [int(r1.iterations), int(r2.iterations)])

    # ---- seam 3: FLIP and APIC steps, particles migrating between the slabs
    for mode, name in ((FLIP, "flip"), (APIC, "apic")):
        hp = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=5))
        hd, hs = jd.hybrid(hp), js.hybrid(hp)
        # a column that sloshes along z so that particles cross the slab faces
        box = ((0.0, 0.0, 0.2), (1.0, 0.6, 0.7))
        v0 = (0.0, 0.0, 1.5)
        nd = hd.emit_box(box[0], box[1], 0.5 / n, v0)
        ns = hs.emit_box(box[0], box[1], 0.5 / n, v0)
        check(name + "_emitted", nd == ns and allsum(np.array([hd.num_local_particles()], dtype=np.float64))[0] == ns,
              [int(nd), int(ns), hd.num_local_particles()])
        moved = 0
        for s in range(steps):
            before = hd.num_local_particles()
            sd, ss = hd.step(1.0 / 120.0), hs.step(1.0 / 120.0)
            moved += abs(hd.num_local_particles() - before)
            check("%s_step%d_pcg" % (name, s), sd.pcg_iterations == ss.pcg_iterations and sd.pcg_last_residual == ss.pcg_last_residual,
                  [int(sd.pcg_iterations), int(ss.pcg_iterations)])
        pd_, vd_ = hd.get_particles()
        ps_, vs_ = hs.get_particles()
        pd_, vd_ = allsum(pd_), allsum(vd_)   # every rank fills only its own particles' entries
        check(name + "_positions_bits", np.array_equal(pd_, ps_), float(np.abs(pd_ - ps_).max()))
        check(name + "_velocities_bits", np.array_equal(vd_, vs_), float(np.abs(vd_ - vs_).max()))
        ugd, ugs = hd.get_velocity(), hs.get_velocity()
        check(name + "_grid_bits", all(np.array_equal(a, b_) for a, b_ in zip(ugd, ugs)))
        check(name + "_sdf_bits", np.array_equal(hd.get_sdf(), hs.get_sdf()))
        check(name + "_pressure_bits", np.array_equal(hd.get_pressure(), hs.get_pressure()))
        ids, lp, lv = hd.get_local_particles()
        z0, z1 = jd.slab_range(n, world, rank)
        ck = np.clip(np.floor(lp[:, 2] * n), 0, n - 1).astype(np.int64) if len(lp) else np.zeros(0, dtype=np.int64)
        total_local = allsum(np.array([len(ids)], dtype=np.float64))[0]
        check(name + "_conservation", total_local == ns, [total_local, int(ns)])
        out[name + "_migrated_net"] = int(moved)
        # round trip through the local set/get pair (particles handed back in a scrambled, un-owned order)
        perm = np.random.default_rng(rank).permutation(len(ids))
        hd.set_local_particles(ids[perm], lp[perm], lv[perm], ns)
        sd2 = hd.step(1.0 / 120.0)
        ss2 = hs.step(1.0 / 120.0)
        p2, _ = hd.get_particles()
        p2 = allsum(p2)
        if mode == FLIP:   # APIC's affine vectors are not part of the local set call: positions only agree for PIC/FLIP
            check(name + "_set_local_roundtrip_bits", np.array_equal(p2, hs.get_particles()[0]))
        hd.close(); hs.close()
    # ---- BASELINE config 5: GridSmokeSolver3 on slabs (MGPCG distributed, advection of the owned planes + all-gather)
    from fluid_engine_dev_b200.jetcu import SmokeParams
    sp = SmokeParams.make(Grid.cube(n), mgpcg=MgpcgParams.make(max_iterations=100, max_levels=5))
    sd_, ss_ = jd.smoke(sp), js.smoke(sp)
    c = (np.arange(n) + 0.5) / n
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
    den = np.ascontiguousarray(np.clip(1.0 - np.sqrt((X - 0.5) ** 2 + (Y - 0.25) ** 2 + (Z - 0.45) ** 2) / 0.2, 0.0, 1.0))
    for q in (sd_, ss_):
        q.set_fields(den, 2.0 * den + 0.1)
    for s in range(steps):
        r1_, r2_ = sd_.step(1.0 / 60.0), ss_.step(1.0 / 60.0)
        check("smoke_step%d_pcg" % s, r1_.iterations == r2_.iterations and r1_.last_residual == r2_.last_residual,
              [int(r1_.iterations), int(r2_.iterations)])
    fd_, fs_ = sd_.get_fields(), ss_.get_fields()
    check("smoke_fields_bits", all(np.array_equal(fd_[k], fs_[k]) for k in fs_))
    check("smoke_moved", float(np.abs(fs_["v"]).max()) > 1e-3)
    sd_.close(); ss_.close()
    msgs, sent = jd.comm_stats()
    out["comm"] = {"messages": msgs, "bytes_sent": sent, "p2p": jd.comm_p2p_enabled()}
    out["ok"] = ok
    outdir = os.environ.get("JETCU_DIST_OUT")
    if outdir:   # one file per rank: concurrent ranks interleave their stdout lines
        with open(os.path.join(outdir, "rank%d.json" % rank), "w") as f:
            json.dump(out, f)
    print(json.dumps(out), flush=True)
    jd.close(); js.close()   # contexts (graphs, then the NCCL communicator) go before torch's process group
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
