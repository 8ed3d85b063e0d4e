"""Per-exchange halo latency of the slab communicator, NCCL send/recv vs the peer-memory mailboxes:
    for p in 0 1; do JETCU_P2P_HALO=$p python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 2953$p scripts/halo_latency.py; done
Rank 0 prints one JSON line: microseconds per exchange (max over ranks) for the plane sizes of the MG levels of a 256^3 domain
and the depths the communication-avoiding smoother uses."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from jetcu_loader import load_package
    load_package()
    from fluid_engine_dev_b200 import Jetcu
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    jc = Jetcu(local)
    jc.comm_init_torch()
    out = {"world": world, "p2p": jc.comm_p2p_enabled(), "us_per_exchange": {}}
    for n in (256, 128, 64, 32):
        for depth in (1, 2, 4, 8):
            if depth > max(1, (n // world) // 4):
                continue
            ms = jc.comm_halo_benchmark(n * n * 8, depth, 200)
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["us_per_exchange"]["%d^2 x %d" % (n, depth)] = round(1e3 * float(t.item()), 2)
    if rank == 0:
        print(json.dumps(out), flush=True)
    jc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
