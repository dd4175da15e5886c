#!/usr/bin/env python3
"""BASELINE.json configs[3] ("C4"): ~10M-triangle city, 8M closest-hit bounce rays + 8M any-hit shadow rays.
BVH memory / L2 pressure study: same ray generators as C2, scene scaled to cells=577 (~560 MB of nodes + Woop
rows, i.e. 4.4x the 126 MB L2). Prints JSON; run under ncu to get L2 hit rate and DRAM bytes."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import adypt_b200 as A
from adypt_b200 import host, workloads as W

def timed(fn, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    cells = int(args[0]) if len(args) > 0 else 577
    reps = int(args[1]) if len(args) > 1 else 5
    mesh = W.city(cells, 1)
    hs = host.HostScene.from_triangles(mesh.positions(), mesh.face_mat, host.materials_array(mesh.materials))
    cache = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), ".cache", "scenes", mesh.name + ".bvh")
    os.makedirs(os.path.dirname(cache), exist_ok=True)
    t0 = time.time()
    if not hs.load_bvh(cache):
        hs.build_bvh(); hs.save_bvh(cache)
    build_s = time.time() - t0
    sc = hs.upload(0)
    tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=7)
    cam = W.city_camera(cells); tr.look(cam['position'], cam['yaw'], cam['pitch'], cam['fov'])
    prim = tr.primary_rays(); ph = sc.trace_closest(prim)
    rays = W.bounce_rays(mesh.positions(), prim, ph['tri'], ph['uv'], per_hit=8, seed=42)
    n = rays.shape[0]
    d_rays = torch.from_numpy(rays).cuda()
    d_tri = torch.empty(n, dtype=torch.int32, device='cuda'); d_t = torch.empty(n, dtype=torch.float32, device='cuda'); d_uv = torch.empty((n, 2), dtype=torch.float32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    ms_closest = timed(lambda: sc.trace_closest(d_rays, d_tri, d_t, d_uv, stream=st), reps)
    stats = sc.trace_stats(d_rays)
    tri = d_tri.cpu().numpy(); uv = d_uv.cpu().numpy()
    shadow = W.shadow_rays(mesh.positions(), tri, uv)
    ns = shadow.shape[0]
    d_sh = torch.from_numpy(shadow).cuda(); d_occ = torch.empty(ns, dtype=torch.uint8, device='cuda')
    ms_any = timed(lambda: sc.trace_any(d_sh, d_occ, stream=st), reps)
    check = {}
    if "--check" in sys.argv:  # parity at C4 size (test infrastructure: the oracle is the checker)
        from oracle import cpu
        woop = cpu.build_woop(hs.tris, hs.tri_indices)
        assert np.array_equal(sc.read_woop().view(np.uint32), woop.view(np.uint32))
        sl = slice(2_000_000, 2_250_000)
        o = cpu.trace_closest(hs.nodes, hs.tri_indices, woop, rays[sl])
        t = d_t.cpu().numpy()
        check = {"oracle_slice_rays": 250000, "ids_equal": float((tri[sl] == o["tri"]).mean()), "t_bit_equal": float((t[sl].view(np.uint32) == o["t"].view(np.uint32)).mean()),
                 "uv_bit_equal": float((uv[sl].view(np.uint32) == o["uv"].view(np.uint32)).all(axis=1).mean()), "oracle_max_stack": o["counters"]["max_stack"]}
        d_any = torch.empty(n, dtype=torch.uint8, device='cuda')
        sc.trace_any(d_rays, d_any, stream=st); torch.cuda.synchronize()
        check["any_equals_closest_hit_all_8M"] = bool(np.array_equal(d_any.cpu().numpy() != 0, tri >= 0))
        oa = cpu.trace_any(hs.nodes, woop, shadow[:250000])
        check["shadow_any_equal"] = float((d_occ.cpu().numpy()[:250000] == oa["occluded"]).mean())
    bpr = 80.0 * stats['nodes'] / n + 48.0 * stats['tris'] / n + 4.0 * stats['hits'] / n + 48.0
    print(json.dumps({"config": f"C4: cells={cells}", "triangles": int(mesh.n_tris), "refs": int(hs.tri_indices.size), "nodes": int(hs.nodes.shape[0]),
                      "bvh_bytes": int(hs.nodes.shape[0] * 80 + hs.tri_indices.size * 52), "scene_device_bytes": sc.device_bytes(), "build_or_load_s": build_s,
                      "closest": {"rays": n, "ms": ms_closest, "Mrays_per_s": n / ms_closest / 1e3, "nodes_per_ray": stats['nodes'] / n, "tris_per_ray": stats['tris'] / n,
                                  "hit_fraction": stats['hits'] / n, "max_stack": stats['max_stack'], "bytes_per_ray": bpr, "algorithmic_GBps": bpr * n / ms_closest / 1e6},
                      "any": {"rays": ns, "ms": ms_any, "Mrays_per_s": ns / ms_any / 1e3, "occluded_fraction": float(d_occ.float().mean().item())}, "parity": check}))

if __name__ == '__main__':
    main()
