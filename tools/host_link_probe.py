"""Host <-> device copy ceiling of the node, measured the way the host-buffer batch call uses it: every rank copies
from / to its own pinned buffers at the same time (torchrun, one rank per GPU), H2D alone, D2H alone and both
directions at once.  The aggregate of the `both` line is the ceiling of `e2e` (bench.py) at that world size.

    python tools/host_link_probe.py [--bind 0|1] [--mb 512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/host_link_probe.py --bind 1

--bind 1 restricts each rank to the cores next to its GPU before the pinned buffers are allocated (vi_b200.numa).
Rank 0 prints one JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", type=int, default=1)
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    from vi_b200 import numa
    dist = None
    torch.cuda.set_device(local)
    binfo = numa.bind_to_gpu(local) if args.bind else {"bound": False}
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.mb << 20
    h_up = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_dn = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_up.fill_(1); h_dn.fill_(2)                       # touch: pages are placed now, under the binding
    d_up = torch.empty(n, dtype=torch.uint8, device=dev)
    d_dn = torch.empty(n, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def run(up, dn):
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.reps if timed else 2):
                if up:
                    with torch.cuda.stream(s_up):
                        d_up.copy_(h_up, non_blocking=True)
                if dn:
                    with torch.cuda.stream(s_dn):
                        h_dn.copy_(d_dn, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        gbs = (int(up) + int(dn)) * n * args.reps / dt / 1e9
        if dist is not None:
            t = torch.tensor([gbs], device=dev, dtype=torch.float64)
            lst = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            per = [float(x.item()) for x in lst]
        else:
            per = [gbs]
        return per

    res = {"h2d": run(True, False), "d2h": run(False, True), "both": run(True, True)}
    if rank == 0:
        out = {"probe": "host_link", "n_gpus": world, "bind": bool(args.bind), "binding_rank0": binfo, "mb": args.mb,
               "aggregate_gbs": {k: round(sum(v), 1) for k, v in res.items()},
               "per_rank_gbs": {k: [round(x, 1) for x in v] for k, v in res.items()}}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
