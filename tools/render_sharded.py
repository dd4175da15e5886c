#!/usr/bin/env python3
"""Sample-sharded progressive render across the GPUs of one box (BASELINE.json configs[4], "C5").

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/render_sharded.py \
           [--width 3840 --height 2160 --spp 1024 --cells 183 --out gpurun_out/c5.exr --check]

One process per GPU. Scene + BVH replicated; tmpLifetime-blocks of sample indices dealt round-robin to ranks;
each rank accumulates a SUM buffer; ONE NCCL all-reduce of W*H*4 floats; rank 0 divides by the sample count and
writes the EXR. --check also renders every sample on rank 0 alone and reports the RMSE between the two images
(different float summation order => tiny, non-zero).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import adypt_b200 as A  # noqa: E402
from adypt_b200 import host, sharding, workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--cells", type=int, default=183)
    ap.add_argument("--out", default="")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    mesh = W.city(args.cells, 1, mixed_materials=True)
    hs = host.HostScene.from_triangles(mesh.positions(), mesh.face_mat, host.materials_array(mesh.materials))
    cache = os.path.join(ROOT, ".cache", "scenes")
    os.makedirs(cache, exist_ok=True)
    bvh_path = os.path.join(cache, mesh.name + ".bvh")
    if rank == 0 and not hs.load_bvh(bvh_path):
        hs.build_bvh()
        hs.save_bvh(bvh_path)
    if world > 1:
        dist.barrier()
    if rank != 0:
        assert hs.load_bvh(bvh_path)
    scene = hs.upload(local)
    tr = A.Tracer(scene, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), args.width, args.height, bias_seed=7)
    cam = W.city_camera(args.cells)
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    ptr, nfl = tr.sum_buffer()
    acc = torch.as_tensor(sharding.DeviceArray(ptr, nfl), device=dev)

    def all_reduce():
        tr.sync()  # the tracer renders on its own stream
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()

    tr.accumulate(0, 16)  # warm-up (allocations, clocks)
    tr.sync()
    if world > 1:
        warm = torch.zeros_like(acc)  # NCCL sets its channels and buffers up at the first collective of a size class
        dist.all_reduce(warm, op=dist.ReduceOp.SUM)
        del warm
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine = sharding.render_sharded(tr, args.spp, rank, world, all_reduce)
    tr.sync()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    out = {"workload": f"C5: {args.width}x{args.height}, {args.spp} spp, sample-sharded over {world} GPU(s), one NCCL all-reduce of {nfl * 4 / 1e6:.1f} MB",
           "n_gpus": world, "seconds": dt, "path_samples_per_s": args.width * args.height * args.spp / dt, "samples_this_rank": mine, "scaling": "strong"}
    if rank == 0:
        img = tr.read(3)
        out["mean"] = float(img.mean())
        if args.out:
            tr.save_exr(args.out, True)
        if args.check:
            tr2 = A.Tracer(scene, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), args.width, args.height, bias_seed=7)
            tr2.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
            tr2.sample(args.spp)
            ref = tr2.read(3)
            out["rmse_vs_single_gpu_running_mean"] = float(np.sqrt(((img - ref) ** 2).mean()))
            out["max_abs_diff"] = float(np.abs(img - ref).max())
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
