#!/usr/bin/env python3
"""Sweep the traversal kernel's tuning knobs on C2 and time the 1080p path tracer (GPU box only)."""
import os, sys, time, json
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
    return min(ts), float(np.median(ts))

def main():
    if os.environ.get('TUNE_LIB'):  # an experimental build from tools/build_variant.py
        A.LIB_PATH = os.path.abspath(os.environ['TUNE_LIB']); print('library', A.LIB_PATH)
    mesh = W.city(183, 1)
    t0 = time.time(); hs = host.build_scene(mesh); print('build %.1fs' % (time.time() - t0))
    sc = hs.upload(0)
    tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=7)
    cam = W.city_camera(183); tr.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    prim = tr.primary_rays(); ph = sc.trace_closest(prim)
    rays = W.bounce_rays(mesh.positions(), prim, ph['tri'], ph['uv'])
    n = rays.shape[0]
    d_rays = torch.from_numpy(rays).cuda(); d_prim = torch.from_numpy(prim).cuda()
    d_tri = torch.empty(n, dtype=torch.int32, device='cuda'); d_t = torch.empty(n, dtype=torch.float32, device='cuda'); d_uv = torch.empty((n, 2), dtype=torch.float32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    base = None
    res = {}
    variants = tuple(int(v) for v in os.environ.get('TUNE_VARIANTS', '0,8').split(','))
    thresholds = tuple(int(v) for v in os.environ.get('TUNE_THRESHOLDS', '24,28,30,32').split(','))
    for variant in variants:
        for thr in thresholds:
            for ctas in (int(os.environ.get('TUNE_CTAS', '0')),):
                sc.configure(ctas, thr, variant)
                mn, md = timed(lambda: sc.trace_closest(d_rays, d_tri, d_t, d_uv, stream=st))
                if base is None: base = d_tri.clone()
                ok = bool(torch.equal(base, d_tri))
                res[(variant, thr)] = mn
                print(f'variant {variant} thr {thr}: min {mn:.3f} ms  med {md:.3f} ms  {n/mn/1e3:.0f} Mrays/s  same={ok}', flush=True)
    best = min(res, key=res.get); print('best', best, res[best])
    sc.configure(0, best[1], best[0])
    m1 = d_tri[:1000000]
    mn, md = timed(lambda: sc.trace_closest(d_prim, d_tri[:1000000], d_t[:1000000], d_uv[:1000000], stream=st))
    print(f'C2-scene 1M primary rays: {mn:.3f} ms {1e3/mn:.0f} Mrays/s')
    occ = torch.empty(n, dtype=torch.uint8, device='cuda')
    for v in variants:
        sc.configure(0, best[1], v)
        mn, md = timed(lambda: sc.trace_any(d_rays, occ, stream=st))
        print(f'any-hit 8M variant {v}: {mn:.3f} ms {n/mn/1e3:.0f} Mrays/s')
    sc.configure(0, best[1], best[0])
    if os.environ.get('TUNE_NO_PT'):
        return
    # path tracer, C3-like
    mesh3 = W.city(183, 1, mixed_materials=True)
    hs3 = host.build_scene(mesh3); sc3 = hs3.upload(0); sc3.configure(0, best[1], best[0])
    for (w, h, spp) in ((1920, 1080, 64),):
        t = A.Tracer(sc3, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), w, h, bias_seed=7)
        t.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
        t.sample(16); t.sync()
        t.trace(False, 0)
        s0 = t.stats()['segments']
        t0 = time.perf_counter(); t.sample(spp); t.sync(); dt = time.perf_counter() - t0
        seg = t.stats()['segments'] - s0
        print(f'PT {w}x{h} {spp}spp: {dt*1e3:.1f} ms  {w*h*spp/dt/1e6:.1f} Msamples/s  {seg/dt/1e6:.0f} Msegments/s  segs/sample {seg/(w*h*spp):.2f}')
        img = t.read(3); print('mean', float(img.mean()))
        t.save_exr('gpurun_out/c3_1080p_64spp.exr', True)

if __name__ == '__main__':
    main()
