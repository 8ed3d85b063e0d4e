"""world_size-2 CPU tests (gloo) of the z-slab host logic — no GPU, no CUDA calls.

What runs here is the DECOMPOSITION PROTOCOL, with the product's own host-side decisions taken from libjetcu.so
(jetcu_slab_plan, jetcu_slab_route_mask, jetcu_slab_range: pure host functions) and the arithmetic of one rank's
kernels restated in numpy in the reference's operation order:
  * red-black SOR on slabs with one halo-plane exchange before each colour pass == the oracle's undivided sweep, bit for bit
    (FdmGaussSeidelSolver3::relaxRedBlack, src/jet/fdm_gauss_seidel_solver3.cpp:150-221);
  * the particle routing rule delivers to every rank exactly the particles of its slab + 2 ghost planes, and the owned
    sets partition the particles.
The CUDA path itself is checked against the single-GPU result in tests/test_gpu_slabs.py (-m gpu)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

LIB = os.path.join(ROOT, "fluid-engine-dev_b200", "libjetcu.so")
GHOST = 2


def _lib():
    lib = C.CDLL(LIB)
    lib.jetcu_slab_route_mask.restype = C.c_uint64
    lib.jetcu_slab_route_mask.argtypes = [C.c_int64, C.c_uint64, C.c_int, C.c_int]
    return lib


def _plan(lib, n, levels, world, rank):
    dl = C.c_uint32(0)
    z0 = (C.c_uint64 * levels)()
    z1 = (C.c_uint64 * levels)()
    rc = lib.jetcu_slab_plan(C.c_uint64(n), C.c_uint64(n), C.c_uint64(n), C.c_uint32(levels), world, rank, C.byref(dl), z0, z1)
    return rc, int(dl.value), [int(v) for v in z0], [int(v) for v in z1]


def _colour_pass(A, b, x, w, colour, k0, k1):
    """One colour pass of relaxRedBlack on planes [k0, k1) of the full-size array x (numpy, reference operation order)."""
    nz, ny, nx = x.shape
    K, J, I = np.meshgrid(np.arange(k0, k1), np.arange(ny), np.arange(nx), indexing="ij")
    sel = ((I + J + K) % 2) == colour
    k, j, i = K[sel], J[sel], I[sel]
    zero = np.zeros(k.shape)

    def term(coef, kk, jj, ii, valid):
        kk2, jj2, ii2 = np.where(valid, kk, 0), np.where(valid, jj, 0), np.where(valid, ii, 0)
        return coef * np.where(valid, x[kk2, jj2, ii2], 0.0), valid

    t = []
    for (coef_idx, dk, dj, di, name) in (((k, j, i - 1), 0, 0, -1, "right"), ((k, j, i), 0, 0, 1, "right"),
                                         ((k, j - 1, i), 0, -1, 0, "up"), ((k, j, i), 0, 1, 0, "up"),
                                         ((k - 1, j, i), -1, 0, 0, "front"), ((k, j, i), 1, 0, 0, "front")):
        kk, jj, ii = k + dk, j + dj, i + di
        valid = (kk >= 0) & (kk < nz) & (jj >= 0) & (jj < ny) & (ii >= 0) & (ii < nx)
        ck, cj, ci = (np.where(valid, c, 0) for c in coef_idx)
        coef = A[name][ck, cj, ci]
        t.append(np.where(valid, coef * x[np.where(valid, kk, 0), np.where(valid, jj, 0), np.where(valid, ii, 0)], zero))
    r = t[0] + t[1] + t[2] + t[3] + t[4] + t[5]
    x[k, j, i] = (1.0 - w) * x[k, j, i] + w * (b[k, j, i] - r) / A["center"][k, j, i]


def _halo(x, k0, k1, rank, world, depth=1):
    """Planes [k0-depth, k0) / [k1, k1+depth) from the neighbouring ranks, exactly what comm_halo_planes does over NCCL."""
    nz = x.shape[0]
    reqs = []
    lo = hi = None
    if rank > 0 and k0 > 0:
        reqs.append(dist.isend(torch.from_numpy(x[k0:k0 + depth].copy()), rank - 1))
        lo = torch.empty((depth,) + x.shape[1:], dtype=torch.float64)
        reqs.append(dist.irecv(lo, rank - 1))
    if rank + 1 < world and k1 < nz:
        reqs.append(dist.isend(torch.from_numpy(x[k1 - depth:k1].copy()), rank + 1))
        hi = torch.empty((depth,) + x.shape[1:], dtype=torch.float64)
        reqs.append(dist.irecv(hi, rank + 1))
    for r in reqs:
        r.wait()
    if lo is not None:
        x[k0 - depth:k0] = lo.numpy()
    if hi is not None:
        x[k1:k1 + depth] = hi.numpy()


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scenes
        from oracle_api import load_port
        lib = _lib()
        res = {}
        # ---- plan: level 0 of 32^3 is split, 16^3 and coarser are replicated; the slabs tile [0, nz)
        n, levels = 32, 4
        rc, dl, z0, z1 = _plan(lib, n, levels, world, rank)
        assert rc == 0 and dl == 1, (rc, dl)
        mine = torch.tensor([z0[0], z1[0]], dtype=torch.int64)
        got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, mine)
        edges = [tuple(int(v) for v in g) for g in got]
        assert edges[0][0] == 0 and edges[-1][1] == n and all(edges[q][1] == edges[q + 1][0] for q in range(world - 1))
        assert all((b - a) % 2 == 0 and b - a >= 4 for a, b in edges)
        assert [z0[l] for l in range(1, levels)] == [0] * (levels - 1) and [z1[l] for l in range(1, levels)] == [16, 8, 4]
        zr0, zr1 = C.c_uint64(0), C.c_uint64(0)
        assert lib.jetcu_slab_range(C.c_uint64(n), world, rank, C.byref(zr0), C.byref(zr1)) == 0
        assert (int(zr0.value), int(zr1.value)) == (z0[0], z1[0])
        assert _plan(lib, 30, 2, world, rank)[0] != 0   # 30 planes cannot be split into even slabs for 2 ranks x 2

        # ---- slab SOR == undivided SOR (oracle), bit for bit
        nn = 16
        A, x0, b = scenes.random_system(nn, nn, nn, seed=3, spd=True)
        k0, k1 = rank * nn // world, (rank + 1) * nn // world
        x = x0.copy()
        sweeps = 3
        for _ in range(sweeps):
            for colour in (0, 1):
                _halo(x, k0, k1, rank, world)
                _colour_pass(A, b, x, 1.5, colour, k0, k1)
        parts = [torch.zeros((nn // world, nn, nn), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(x[k0:k1].copy()))
        full = np.concatenate([p.numpy() for p in parts], axis=0)
        oracle = load_port()
        want = oracle.relax_red_black(A, b, x0, 1.5, sweeps)
        assert np.array_equal(full, want), float(np.abs(full - want).max())

        # ---- communication-avoiding form (launch_relax in csrc/fdm.cu): ONE exchange of 2k planes per block of k sweeps,
        # the halo region recomputed redundantly on a window shrinking by one plane per colour pass — same bits
        # a rank's copy of x and b is only meaningful on the planes it owns: poison the rest, so that any read of a
        # plane that was not exchanged shows up as NaN (b is read-only for the sweeps: one exchange of 2*kmax-1 planes)
        x = x0.copy()
        bb = b.copy()
        x[:k0] = np.nan; x[k1:] = np.nan
        bb[:k0] = np.nan; bb[k1:] = np.nan
        _halo(bb, k0, k1, rank, world, depth=2 * 2 - 1)
        b_full, b = b, bb
        done = 0
        for k in (2, 1):                      # 3 sweeps as a block of 2 and a block of 1
            _halo(x, k0, k1, rank, world, depth=2 * k)
            for s2 in range(k):
                for colour in (0, 1):
                    ext = 2 * (k - s2) - 1 - colour
                    _colour_pass(A, b, x, 1.5, colour, max(k0 - ext, 0), min(k1 + ext, nn))
            done += k
        assert done == sweeps
        b = b_full
        dist.all_gather(parts, torch.from_numpy(x[k0:k1].copy()))
        full = np.concatenate([p.numpy() for p in parts], axis=0)
        assert np.array_equal(full, want), float(np.abs(full - want).max())

        # ---- particle routing: every rank ends up with exactly its extended slab; owners partition the set
        nzp = 24
        T = nzp // world
        rng = np.random.default_rng(100 + rank)
        planes = rng.integers(0, nzp, size=500)             # particles currently HELD by this rank, anywhere in z
        ids = rank * 1000 + np.arange(500)
        send = {q: [] for q in range(world)}
        keep = []
        for pid, ck in zip(ids, planes):
            m = int(lib.jetcu_slab_route_mask(int(ck), nzp, world, rank))
            for q in range(world):
                if m >> q & 1:
                    assert q != rank
                    send[q].append((int(pid), int(ck)))
            if rank * T - GHOST <= ck < (rank + 1) * T + GHOST:
                keep.append((int(pid), int(ck)))
        boxes = [None] * world
        dist.all_gather_object(boxes, send)
        have = keep + [pc for q in range(world) if q != rank for pc in boxes[q][rank]]
        allp = [None] * world
        dist.all_gather_object(allp, list(zip(ids.tolist(), planes.tolist())))
        everything = [pc for lst in allp for pc in lst]
        want_ext = sorted(pc for pc in everything if rank * T - GHOST <= pc[1] < (rank + 1) * T + GHOST)
        assert sorted(have) == want_ext
        owned = [pc for pc in have if rank * T <= pc[1] < (rank + 1) * T]
        counts = [None] * world
        dist.all_gather_object(counts, len(owned))
        assert sum(counts) == len(everything)
        res["ok"] = True
        with open(os.path.join(tmp, "rank%d.ok" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_slab_protocol_world2(tmp_path):
    if not os.path.exists(LIB):
        pytest.fail("libjetcu.so not built (python fluid-engine-dev_b200/build.py)")
    world = 2
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "rank%d.ok" % r)) for r in range(world))


def test_slab_plan_shapes():
    """The plan at the benchmark sizes: 256^3 on 2/4/8 ranks and 512^3 on 8."""
    if not os.path.exists(LIB):
        pytest.fail("libjetcu.so not built")
    lib = _lib()
    for n, world, want_dist in ((256, 2, 2), (256, 4, 2), (256, 8, 2), (512, 8, 3), (64, 2, 1), (64, 8, 1)):   # levels of 64^3 and below are replicated
        levels = 6
        for rank in range(world):
            rc, dl, z0, z1 = _plan(lib, n, levels, world, rank)
            assert rc == 0 and dl == want_dist, (n, world, rank, rc, dl)
            for l in range(levels):
                nz = n >> l
                if l < dl:
                    assert (z0[l], z1[l]) == (rank * nz // world, (rank + 1) * nz // world) and (z1[l] - z0[l]) % 2 == 0
                else:
                    assert (z0[l], z1[l]) == (0, nz)
    # routing: a particle deep inside a slab goes nowhere; one in the two planes next to a face is copied to the neighbour
    assert lib.jetcu_slab_route_mask(10, 64, 2, 0) == 0
    assert lib.jetcu_slab_route_mask(31, 64, 2, 0) == 0b10 and lib.jetcu_slab_route_mask(30, 64, 2, 0) == 0b10
    assert lib.jetcu_slab_route_mask(29, 64, 2, 0) == 0
    assert lib.jetcu_slab_route_mask(40, 64, 2, 0) == 0b10      # migrated into rank 1's slab
    assert lib.jetcu_slab_route_mask(32, 64, 4, 1) == 0b100     # left rank 1 for rank 2 (and stays rank 1's ghost)
