#!/usr/bin/env python3
"""Host <-> device copy ceiling of the box with N ranks copying AT THE SAME TIME (one rank per GPU under torchrun): what the
end-to-end leg of bench.py (adypt_trace_closest on pinned host arrays: 256 MB up + 128 MB down per GPU and step) can reach at
most at N GPUs. No kernels, no product code: pinned cudaMemcpyAsync only.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe_multi.py

Prints one JSON line on rank 0: per-direction and duplex GB/s per GPU and aggregate, and the Mrays/s ceiling they imply."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 8_000_000
h_in = torch.empty(n * 8, dtype=torch.float32).pin_memory()
h_out = torch.empty(n * 4, dtype=torch.float32).pin_memory()
h_in.fill_(1.0)
d_in = torch.empty(n * 8, dtype=torch.float32, device=dev)
d_out = torch.empty(n * 4, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def up():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    up()
    down()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
    mn = dt.clone()
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    barrier()
    return float(dt.item()), float(mn.item())


a, a0 = timed(up)
b, b0 = timed(down)
c, c0 = timed(both)
if rank == 0:
    print(json.dumps({
        "n_gpus": world, "bytes_up": n * 32, "bytes_down": n * 16,
        "h2d_ms_slowest_rank": a * 1e3, "h2d_ms_fastest_rank": a0 * 1e3, "h2d_GBps_per_gpu": n * 32 / a / 1e9, "h2d_GBps_aggregate": world * n * 32 / a / 1e9,
        "d2h_ms_slowest_rank": b * 1e3, "d2h_ms_fastest_rank": b0 * 1e3, "d2h_GBps_per_gpu": n * 16 / b / 1e9, "d2h_GBps_aggregate": world * n * 16 / b / 1e9,
        "duplex_ms_slowest_rank": c * 1e3, "duplex_ms_fastest_rank": c0 * 1e3, "duplex_GBps_aggregate": world * n * 48 / c / 1e9,
        "e2e_ceiling_Mrays_per_s": world * n / c / 1e6,
        "how": "every rank copies its own pinned 256 MB up / 128 MB down (and both at once on two streams) 20 times between barriers; times are the slowest rank's mean",
    }), flush=True)
if world > 1:
    dist.destroy_process_group()
