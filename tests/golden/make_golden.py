"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libjetref.so, which oracle/Makefile
compiles from /root/reference).  Run here (the container that has /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

The GPU box has no /root/reference; the committed .npz files carry the reference's answers there.
Inputs are regenerated from seeds by tests/scenes.py; outputs (and the few inputs that are cheap) are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import scenes  # noqa: E402
from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams, load_ref  # noqa: E402


def main():
    ref = load_ref()
    assert ref.kind == "reference"
    ref.set_num_threads(0)

    # 1. linear building blocks on a random 7-point system
    A, x, b = scenes.random_system(10, 8, 8, seed=11)
    f = np.random.default_rng(12).uniform(-1, 1, (8, 8, 12))
    c = np.random.default_rng(13).uniform(-1, 1, (4, 4, 6))
    np.savez_compressed(os.path.join(HERE, "linear_ops.npz"), mvm=ref.mvm(A, x), residual=ref.residual(A, x, b),
                        relax2=ref.relax_red_black(A, b, x, 1.5, 2), restrict=ref.restrict(f),
                        correct=ref.correct(c, f), dot=np.array(ref.dot(x, b)), axpy=ref.axpy(-0.7, x, b))

    # 2. V-cycle + MGPCG on the Neumann Poisson fixture
    n, levels = 16, 3
    Al = scenes.poisson_levels(n, levels)
    rng = np.random.default_rng(3)
    bb = rng.uniform(-1, 1, (n, n, n)); bb -= bb.mean()
    x0 = rng.uniform(-1, 1, (n, n, n))
    p = MgpcgParams.make(max_iterations=25, max_levels=levels)
    xs, res = ref.mgpcg_solve(Al, bb, p)
    np.savez_compressed(os.path.join(HERE, "mgpcg_poisson16.npz"), vcycle=ref.vcycle(Al, bb, x0, p), x=xs,
                        iterations=np.array(res.iterations), last_residual=np.array(res.last_residual))

    # 3. pressure projection, dam column + box collider, 16^3 and 32^3
    for n, levels in ((16, 3), (32, 5)):
        sc = scenes.projection_scene(n, "column", True)
        p = MgpcgParams.make(max_iterations=200, max_levels=levels)
        u, v, w, pr, res = ref.pressure_solve(sc["grid"], sc["u"], sc["v"], sc["w"], p, collider_sdf=sc["collider_sdf"],
                                              fluid_sdf=sc["fluid_sdf"])
        s0 = ref.pressure_get_system(sc["grid"], 0)
        sl = ref.pressure_get_system(sc["grid"], levels - 1)
        np.savez_compressed(os.path.join(HERE, "pressure_column_box_%d.npz" % n), u=u, v=v, w=w, p=pr,
                            iterations=np.array(res.iterations), last_residual=np.array(res.last_residual),
                            true_residual=np.array(res.true_residual_l2), b_l2=np.array(res.b_l2),
                            A0=s0["A"], b0=s0["b"], sdf0=s0["fluid_sdf"], uw0=s0["u_weights"], vw0=s0["v_weights"],
                            ww0=s0["w_weights"], A_last=sl["A"], uw_last=sl["u_weights"], sdf_last=sl["fluid_sdf"])

    # 4. hash-grid binning
    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.3, 1.3, (3000, 3))
    r = ref.bin_points(pts, (8, 8, 8), 0.11)
    bucket, rank = ref.bin_points_serial(pts, (8, 8, 8), 0.11)
    np.savez_compressed(os.path.join(HERE, "bin_points.npz"), keys_unsorted=r["keys_unsorted"],
                        keys_sorted=r["keys_sorted"], start_table=r["start_table"], end_table=r["end_table"],
                        serial_rank=rank)

    # 5. PIC / FLIP / APIC: 5 sub-steps of an 8^3 dam break (stage outputs of the first step + final particles)
    for mode, name in ((PIC, "pic"), (FLIP, "flip"), (APIC, "apic")):
        n = 8
        prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=2))
        h = ref.hybrid(prm)
        h.set_particles(scenes.dam_particles(n))
        dt = 1 / 60.0
        out = {}
        h.begin_advance_time_step(dt)
        out["u_begin"], out["v_begin"], out["w_begin"] = h.get_velocity()
        out["sdf"] = h.get_sdf()
        out["um"], out["vm"], out["wm"] = h.get_markers()
        h.compute_external_forces(dt)
        r0 = h.compute_pressure(dt)
        out["u_press"], out["v_press"], out["w_press"] = h.get_velocity()
        out["pressure"] = h.get_pressure()
        out["iterations0"] = np.array(r0.iterations)
        h.compute_advection(dt)
        out["pos1"], out["vel1"] = h.get_particles()
        iters = [r0.iterations]
        for _ in range(4):
            iters.append(h.step(dt).pcg_iterations)
        out["pos5"], out["vel5"] = h.get_particles()
        out["iterations"] = np.array(iters)
        if mode == APIC:
            out["cx5"], out["cy5"], out["cz5"] = h.get_affine()
        np.savez_compressed(os.path.join(HERE, "hybrid_%s_8.npz" % name), **out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
