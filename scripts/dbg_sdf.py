import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
from gpu_common import checker, jetcu_or_fail
from oracle_api import APIC, FLIP, PIC, Grid, HybridParams, MgpcgParams
jc = jetcu_or_fail(); oracle = checker()
mode, n, levels = PIC, 16, 3
prm = HybridParams.make(Grid.cube(n), mode=mode, mgpcg=MgpcgParams.make(max_iterations=100, max_levels=levels))
a, b = jc.hybrid(prm), oracle.hybrid(prm)
pts = scenes.dam_particles(n)
perm = np.random.default_rng(1).permutation(len(pts))
pts = np.ascontiguousarray(pts[perm])
vel = np.random.default_rng(2).uniform(-0.2, 0.2, pts.shape)
a.set_particles(pts, vel); b.set_particles(pts, vel)
dt = 1/60.
for step in range(3):
    a.begin_advance_time_step(dt); b.begin_advance_time_step(dt)
    sa, sb = a.get_sdf(), b.get_sdf()
    bad = np.argwhere(sa != sb)
    print("step", step, "mismatch", len(bad), "of", sa.size)
    pa, _ = a.get_particles(); pb, _ = b.get_particles()
    print("  particles equal:", np.array_equal(pa, pb), np.abs(pa-pb).max())
    for (k, j, i) in bad[:8]:
        c = (np.array([i, j, k]) + 0.5) / n
        d = np.sqrt(((pb - c) ** 2).sum(1)); dg = np.sqrt(((pa - c) ** 2).sum(1))
        print("  cell", i, j, k, "gpu", sa[k, j, i], "ref", sb[k, j, i], "diff", sa[k,j,i]-sb[k,j,i], "brute(ref parts)", min(d.min(), 2*1.2/n/np.sqrt(2)) - 1.2/n/np.sqrt(2),
              "brute(gpu parts)", min(dg.min(), 2*1.2/n/np.sqrt(2)) - 1.2/n/np.sqrt(2))
    a.compute_external_forces(dt); b.compute_external_forces(dt)
    a.compute_pressure(dt); b.compute_pressure(dt)
    a.compute_advection(dt); b.compute_advection(dt)
