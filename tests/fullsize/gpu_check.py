#!/usr/bin/env python3
"""First-light check on a GPU box: product (CUDA) vs oracle (CPU) on C1, a C2 slice and a small render."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adypt_b200 as A
from adypt_b200 import workloads as W
from oracle import ref, cpu

def scene_for(mesh):
    b = ref.build(mesh.write_obj('.cache/scenes'))
    return b

def main():
    print('devices', A.device_count())
    mesh = W.sphere_lattice(5); b = scene_for(mesh)
    sc = A.Scene(b.nodes, b.tri_indices, None, b.tris, b.mats)
    w = sc.read_woop()
    print('woop bit-exact vs reference:', np.array_equal(w.view(np.uint32), b.woop.view(np.uint32)))
    cfg = A.PTConfig.make()
    tr = A.Tracer(sc, cfg, 1000, 1000, bias_seed=7)
    cam = W.lattice_camera()
    tr.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    rays = tr.primary_rays()
    m = cpu.camera_matrices(cam['fov'], cam['yaw'], cam['pitch'], 1000, 1000)
    orays = cpu.primary_rays(cam['position'], 1e-4, m['inv_proj'], m['inv_view'], 1000, 1000)
    print('primary rays bit-exact:', np.array_equal(rays.view(np.uint32), orays.view(np.uint32)))
    t0 = time.time(); g = sc.trace_closest(rays); t1 = time.time()
    o = cpu.trace_closest(b.nodes, b.tri_indices, b.woop, rays)
    print('C1 ids equal:', np.array_equal(g['tri'], o['tri']), 't equal:', np.array_equal(g['t'].view(np.uint32), o['t'].view(np.uint32)),
          'uv equal:', np.array_equal(g['uv'].view(np.uint32), o['uv'].view(np.uint32)), 'host-call s', t1 - t0)
    # small render
    tr2 = A.Tracer(sc, cfg, 256, 192, bias_seed=7)
    tr2.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    tr2.sample(32)
    img = tr2.read()
    m2 = cpu.camera_matrices(cam['fov'], cam['yaw'], cam['pitch'], 256, 192)
    oc = dict(max_bounce=5, subpixel=8, tmp_lifetime=16, ray_tmin=1e-4, clamp=4.0, sun=(1.0, 1.0, 1.0))
    oimg, _, cnt = cpu.pt_render(b, cam['position'], m2['inv_proj'], m2['inv_view'], 256, 192, oc, tr2.get_bias(), 0, 32)
    oimg = oimg.reshape(192, 256, 4)
    d = img[..., :3] - oimg[..., :3]
    print('render rmse', float(np.sqrt((d ** 2).mean())), 'max', float(np.abs(d).max()), 'mean img', float(img[..., :3].mean()),
          'exact pixels', float((np.abs(d).max(axis=2) == 0).mean()), tr2.stats(), cnt)
    # C2
    mesh = W.city(183, 1); b = scene_for(mesh)
    sc2 = A.Scene(b.nodes, b.tri_indices, b.woop, b.tris, b.mats)
    cam = W.city_camera(183)
    tr3 = A.Tracer(sc2, cfg, 1000, 1000, bias_seed=7)
    tr3.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    rays = tr3.primary_rays()
    g = sc2.trace_closest(rays)
    br = W.bounce_rays(mesh.positions(), rays, g['tri'], g['uv'])
    print('bounce rays', br.shape)
    import torch
    d_rays = torch.from_numpy(br).cuda()
    n = br.shape[0]
    d_tri = torch.empty(n, dtype=torch.int32, device='cuda'); d_t = torch.empty(n, dtype=torch.float32, device='cuda'); d_uv = torch.empty((n, 2), dtype=torch.float32, device='cuda')
    for thr in (32, 28, 24, 16, 8, 1):
        sc2.configure(0, thr)
        for it in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); sc2.trace_closest(d_rays, d_tri, d_t, d_uv, stream=torch.cuda.current_stream().cuda_stream); e1.record(); torch.cuda.synchronize()
        print('thr', thr, 'C2 8M incoherent: %.3f ms  %.1f Mrays/s' % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e3))
    sub = slice(0, 2000000)
    o = cpu.trace_closest(b.nodes, b.tri_indices, b.woop, br[sub])
    gt = d_tri.cpu().numpy()[sub]; gtt = d_t.cpu().numpy()[sub]
    print('C2 ids equal frac', float((gt == o['tri']).mean()), 't bit-equal frac', float((gtt.view(np.uint32) == o['t'].view(np.uint32)).mean()))
    occ = torch.empty(n, dtype=torch.uint8, device='cuda')
    sc2.trace_any(d_rays, occ, stream=torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    oa = cpu.trace_any(b.nodes, b.woop, br[sub])
    print('any equal frac', float((occ.cpu().numpy()[sub] == oa['occluded']).mean()))

if __name__ == '__main__':
    main()
