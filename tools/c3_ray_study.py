#!/usr/bin/env python3
"""Why do C3's bounce rays trace at ~3.0 Grays/s when C2's trace at 4.5? Host-made diffuse bounce rays from the 1920x1080
view (the C3 camera) against the 1000x1000 view (C2), same city, with node / triangle counts per ray (GPU box only)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adypt_b200 as A
from adypt_b200 import workloads as W, host

def timed(fn, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

def main():
    mesh = W.city(183, 1)
    sc = host.build_scene(mesh).upload(0)
    cam = W.city_camera(183)
    st = torch.cuda.current_stream().cuda_stream
    for (w, h, per_hit) in ((1000, 1000, 8), (1920, 1080, 4), (1920, 1080, 8)):
        tr = A.Tracer(sc, A.PTConfig.make(), w, h, bias_seed=7)
        tr.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
        prim = tr.primary_rays(); ph = sc.trace_closest(prim)
        rays = W.bounce_rays(mesh.positions(), prim, ph['tri'], ph['uv'], per_hit=per_hit)
        n = rays.shape[0]
        d = torch.from_numpy(rays).cuda()
        tri = torch.empty(n, dtype=torch.int32, device='cuda'); uv = torch.empty((n, 2), dtype=torch.float32, device='cuda')
        ms = timed(lambda: sc.trace_closest(d, tri, None, uv, stream=st))
        s = sc.trace_stats(d)
        print(f'{w}x{h} x{per_hit}: {n} rays {ms:.3f} ms {n/ms/1e3:.0f} Mrays/s  nodes/ray {s["nodes"]/n:.2f} tris/ray {s["tris"]/n:.2f} hit {s["hits"]/n:.3f} primary-hit {float((ph["tri"]>=0).mean()):.3f}', flush=True)
        # the same rays in pixel-major-per-sample order (what the wavefront queue holds): sample k of every pixel, then k+1
        perm = np.arange(n).reshape(-1, per_hit).T.reshape(-1)
        d2 = torch.from_numpy(rays[perm]).cuda()
        ms = timed(lambda: sc.trace_closest(d2, tri, None, uv, stream=st))
        print(f'   sample-major order: {ms:.3f} ms {n/ms/1e3:.0f} Mrays/s', flush=True)
        del tr, d, d2, tri, uv

if __name__ == '__main__':
    main()
