"""Multi-rank host logic on CPU: sample-block sharding + all-reduce of the sum accumulator, world_size 2 over
gloo. The per-rank renderer here is the CPU oracle in sum mode (tests may use it); the product's
adypt_b200.sharding.render_sharded drives it exactly as it drives the CUDA tracer on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adypt_b200 import sharding
from conftest import load_golden

W, H, SPP = 24, 16, 40
CFG = dict(max_bounce=4, subpixel=4, tmp_lifetime=8, ray_tmin=1e-4, clamp=4.0, sun=(1.0, 0.9, 0.8))


def test_block_partition_covers_every_sample_once():
    for total, L, world in [(64, 16, 8), (40, 8, 2), (1024, 16, 8), (17, 16, 4), (5, 16, 2)]:
        seen = []
        for r in range(world):
            for first, n in sharding.blocks_for_rank(total, L, r, world):
                assert first % L == 0 and 0 < n <= L
                seen += list(range(first, first + n))
        assert sorted(seen) == list(range(total))
    assert sharding.sample_blocks(40, 16) == [(0, 16), (16, 16), (32, 8)]


def test_ray_ranges_are_contiguous_and_balanced():
    for n, world in [(8_000_000, 8), (10, 4), (3, 8), (0, 2)]:
        ranges = [sharding.ray_range_for_rank(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1


class OracleTracer:
    """CPU stand-in with the tracer interface render_sharded needs."""

    def __init__(self, g, cam, bias):
        from oracle import cpu
        self.cpu, self.g, self.bias = cpu, g, bias
        self.tmp_lifetime = CFG["tmp_lifetime"]
        self.m = cpu.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), W, H)
        self.origin = cam[:3]
        self.sum = np.zeros((W * H, 4), dtype=np.float32)
        self.result = None

    def clear_sum(self):
        self.sum[:] = 0

    def accumulate(self, first, n):
        self.cpu.pt_render(self.g, self.origin, self.m["inv_proj"], self.m["inv_view"], W, H, CFG, self.bias, first, n,
                           out_rgba=self.sum, nthreads=2, sum_mode=True)

    def resolve_sum(self):
        self.result = self.sum[:, :3] / self.sum[:, 3:4]


def _bias():
    return np.random.default_rng(7).integers(0, 256, size=W * H * 2, dtype=np.uint8)


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("city12")
    tr = OracleTracer(g, g.extra["cam"], _bias())

    def all_reduce():
        t = torch.from_numpy(tr.sum)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    previews = []

    def preview(blocks_done):
        t = torch.from_numpy(tr.sum.copy())  # reduce a COPY: the accumulator keeps growing
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        previews.append((blocks_done, float(t[:, 3].min()), float(t[:, 3].max())))

    mine = sharding.render_sharded(tr, SPP, rank, world, all_reduce, preview_every=1, preview=preview)
    # 5 blocks over 2 ranks: rank 0 has 3, rank 1 has 2 -> previews after rounds 1 and 2 with 16 and 32 samples/pixel
    assert [p[0] for p in previews] == [1, 2] and previews[0][1:] == (16.0, 16.0) and previews[1][1:] == (32.0, 32.0)
    counts = torch.tensor([mine])
    dist.all_reduce(counts)
    if rank == 0:
        np.savez(out_path, img=tr.result, total=counts.numpy(), w=tr.sum[:, 3])
    dist.destroy_process_group()


def test_two_rank_render_equals_single_rank(tmp_path, cpu):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    assert int(z["total"][0]) == SPP and np.all(z["w"] == SPP)
    g = load_golden("city12")
    single = OracleTracer(g, g.extra["cam"], _bias())
    sharding.render_sharded(single, SPP, 0, 1, lambda: None)
    # same samples, different float summation order across ranks => tiny differences only
    assert np.allclose(z["img"], single.result, rtol=0, atol=2e-6 * CFG["clamp"])
    # and the sum/divide image equals the reference's running mean up to rounding
    mean, _, _ = cpu.pt_render(g, single.origin, single.m["inv_proj"], single.m["inv_view"], W, H, CFG, _bias(), 0, SPP, nthreads=2)
    assert np.sqrt(((mean[:, :3] - single.result) ** 2).mean()) < 1e-6
