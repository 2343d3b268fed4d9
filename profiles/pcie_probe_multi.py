#!/usr/bin/env python3
"""Aggregate pinned-copy bandwidth of the HOST with all N GPUs copying at once (the ceiling of every end-to-end figure at N > 1).
Launch like bench.py:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/pcie_probe_multi.py [GiB]
Every rank copies `GiB` per direction on its own GPU between barriers; rank 0 prints one JSON line with the per-rank and
summed rates for D2H alone, H2D alone and both directions at once (wall clock between barriers, max over ranks)."""
import json, os, sys, time
import torch
import torch.distributed as dist

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(gib * (1 << 30))
d, d2 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
h, h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


def d2h():
    h.copy_(d, non_blocking=True)


def h2d():
    d2.copy_(h2, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


for f in (d2h, h2d, both):
    f()
t1, t2, t3 = timed(d2h), timed(h2d), timed(both)
if rank == 0:
    print(json.dumps({"n_gpus": world, "gib_per_gpu_per_direction": gib, "cpus": os.cpu_count(),
                      "d2h_gbs_per_gpu": n / t1 / 1e9, "h2d_gbs_per_gpu": n / t2 / 1e9, "duplex_each_gbs_per_gpu": n / t3 / 1e9,
                      "d2h_gbs_total": world * n / t1 / 1e9, "h2d_gbs_total": world * n / t2 / 1e9, "duplex_each_gbs_total": world * n / t3 / 1e9}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
