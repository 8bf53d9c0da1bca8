#!/usr/bin/env python
"""Aggregate host->device copy bandwidth of the box with N ranks copying at once (torchrun): the ceiling of the
host-buffer (e2e) numbers of bench.py at N GPUs.  Each rank copies a 1 GiB pinned buffer to its GPU `iters` times
between barriers; rank 0 prints per-rank and aggregate GB/s.  Measurement tool only."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 1 << 30
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(rank + 1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = {}
    for mode in ("solo", "all"):
        rates = []
        for who in (range(world) if mode == "solo" else [None]):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            iters = 8
            t0 = time.perf_counter()
            if who is None or who == rank:
                for _ in range(iters):
                    d.copy_(h, non_blocking=True)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            mine = iters * nbytes / dt / 1e9 if (who is None or who == rank) else 0.0
            t = torch.tensor([mine], dtype=torch.float64, device=dev)
            if world > 1:
                g = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(g, t)
                vals = [float(x.item()) for x in g]
            else:
                vals = [mine]
            rates.append(vals)
        if mode == "solo":
            res["solo_gbs_per_rank"] = [rates[r][r] for r in range(world)]
        else:
            res["concurrent_gbs_per_rank"] = rates[0]
            res["concurrent_gbs_total"] = sum(rates[0])
    if rank == 0:
        res["n_gpus"] = world
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
