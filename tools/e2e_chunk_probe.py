"""Times adypt_trace_closest on pinned HOST arrays (the e2e leg of bench.py) for the current ADYPT_HOST_CHUNK (GPU box only)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adypt_b200 as A
from adypt_b200 import host, workloads as W
mesh = W.city(183, 1); hs = host.build_scene(mesh); sc = hs.upload(0)
tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=7)
cam = W.city_camera(183); tr.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
prim = tr.primary_rays(); ph = sc.trace_closest(prim)
rays = W.bounce_rays(mesh.positions(), prim, ph['tri'], ph['uv']); n = rays.shape[0]
h_rays = torch.from_numpy(rays).pin_memory(); h_tri = torch.empty(n, dtype=torch.int32).pin_memory(); h_t = torch.empty(n, dtype=torch.float32).pin_memory(); h_uv = torch.empty((n, 2), dtype=torch.float32).pin_memory()
for _ in range(3): sc.trace_closest(h_rays, h_tri, h_t, h_uv)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); sc.trace_closest(h_rays, h_tri, h_t, h_uv); ts.append(time.perf_counter() - t0)
print('chunk', os.environ.get('ADYPT_HOST_CHUNK'), 'e2e ms min %.3f med %.3f -> %.0f Mrays/s' % (min(ts)*1e3, np.median(ts)*1e3, n/np.median(ts)/1e6))
