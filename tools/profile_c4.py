#!/usr/bin/env python3
"""Runs the closest-hit kernel on C4 (10M-triangle city, 8M incoherent rays) a few times (for ncu: -k regex:trace_kernel -s 3 -c 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adypt_b200 as A
from adypt_b200 import host, workloads as W
mesh = W.city(577, 1)
sc = host.build_scene(mesh).upload(0)
tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=7)
cam = W.city_camera(577); tr.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
prim = tr.primary_rays(); ph = sc.trace_closest(prim)
rays = W.bounce_rays(mesh.positions(), prim, ph['tri'], ph['uv'])
n = rays.shape[0]
d_rays = torch.from_numpy(rays).cuda()
d_tri = torch.empty(n, dtype=torch.int32, device='cuda'); d_t = torch.empty(n, dtype=torch.float32, device='cuda'); d_uv = torch.empty((n, 2), dtype=torch.float32, device='cuda')
for _ in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); sc.trace_closest(d_rays, d_tri, d_t, d_uv, stream=torch.cuda.current_stream().cuda_stream); e1.record(); torch.cuda.synchronize()
    print('C4 closest ms', e0.elapsed_time(e1))
