#!/usr/bin/env python3
"""Does re-ordering incoherent rays pay? Times the closest-hit kernel on the C2 ray set in spawn order, sorted by
direction octant, by origin Morton code, and by (octant, Morton). GPU box only; prints one line per ordering."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adypt_b200 as A
from adypt_b200 import host, workloads as W

def morton3(q):
    def spread(v):
        v = v.astype(np.uint64) & 0x1fffff
        v = (v | (v << 32)) & 0x1f00000000ffff
        v = (v | (v << 16)) & 0x1f0000ff0000ff
        v = (v | (v << 8)) & 0x100f00f00f00f00f
        v = (v | (v << 4)) & 0x10c30c30c30c30c3
        v = (v | (v << 2)) & 0x1249249249249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)

def main():
    mesh = W.city(183, 1); hs = host.build_scene(mesh); sc = hs.upload(0)
    tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=7)
    cam = W.city_camera(183); tr.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    prim = tr.primary_rays(); ph = sc.trace_closest(prim)
    rays = W.bounce_rays(mesh.positions(), prim, ph['tri'], ph['uv']); n = rays.shape[0]
    octant = ((rays[:, 4] < 0).astype(np.uint64) | ((rays[:, 5] < 0).astype(np.uint64) << 1) | ((rays[:, 6] < 0).astype(np.uint64) << 2))
    lo, hi = rays[:, :3].min(0), rays[:, :3].max(0)
    q = np.clip(((rays[:, :3] - lo) / (hi - lo) * 1023.0), 0, 1023).astype(np.uint64)
    mort = morton3(q)
    orders = {"spawn order": np.arange(n), "by octant": np.argsort(octant, kind="stable"), "by origin morton": np.argsort(mort, kind="stable"),
              "by octant, morton": np.argsort((octant << np.uint64(60)) | mort, kind="stable"), "random": np.random.default_rng(1).permutation(n)}
    st = torch.cuda.current_stream().cuda_stream
    d_tri = torch.empty(n, dtype=torch.int32, device='cuda'); d_t = torch.empty(n, dtype=torch.float32, device='cuda'); d_uv = torch.empty((n, 2), dtype=torch.float32, device='cuda')
    base = None
    for name, perm in orders.items():
        d_rays = torch.from_numpy(np.ascontiguousarray(rays[perm])).cuda()
        ts = []
        for _ in range(6):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); sc.trace_closest(d_rays, d_tri, d_t, d_uv, stream=st); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        tri = np.empty(n, dtype=np.int32); tri[perm] = d_tri.cpu().numpy()
        if base is None: base = tri
        stt = sc.trace_stats(d_rays)
        print(f"{name:20s} {min(ts[1:]):.3f} ms  {n/min(ts[1:])/1e3:.0f} Mrays/s  same={np.array_equal(base, tri)}", flush=True)

if __name__ == "__main__":
    main()
