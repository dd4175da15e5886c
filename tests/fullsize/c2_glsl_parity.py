#!/usr/bin/env python3
"""BASELINE.json configs[1] ("C2") at FULL size against the reference's own traversal.glsl run on the CPU
(oracle/_ref/libadypt_glsl.so): all 8 000 000 incoherent rays, closest-hit ids + uv and any-hit bits. Prints JSON.
Test infrastructure: uses oracle/ as the checker."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adypt_b200 as A
from adypt_b200 import host, workloads as W
from oracle import cpu, glsl_ref

def main():
    mesh = W.city(183, 1)
    hs = host.build_scene(mesh)
    sc = hs.upload(0)
    tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=7)
    cam = W.city_camera(183); tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    prim = tr.primary_rays(); ph = sc.trace_closest(prim)
    rays = W.bounce_rays(mesh.positions(), prim, ph["tri"], ph["uv"], per_hit=8, seed=42)
    g = sc.trace_closest(rays)
    occ = sc.trace_any(rays)
    woop = cpu.build_woop(hs.tris, hs.tri_indices)
    t0 = time.perf_counter(); tri, uv = glsl_ref.trace_closest(hs.nodes, hs.tri_indices, woop, rays); s_closest = time.perf_counter() - t0
    t0 = time.perf_counter(); rocc = glsl_ref.trace_any(hs.nodes, woop, rays); s_any = time.perf_counter() - t0
    ptri, puv = glsl_ref.trace_closest(hs.nodes, hs.tri_indices, woop, prim)
    hit = tri >= 0
    print(json.dumps({"config": f"C2: {mesh.n_tris} triangles, {rays.shape[0]} incoherent rays + {prim.shape[0]} primary rays", "hit_fraction": float(hit.mean()),
                      "ids_equal_reference_glsl": bool(np.array_equal(g["tri"], tri)),
                      "uv_bits_equal_on_hits": bool(np.array_equal(g["uv"].view(np.uint32)[hit], uv.view(np.uint32)[hit])),
                      "any_hit_equal_reference_glsl": bool(np.array_equal(occ, rocc)),
                      "primary_ids_equal": bool(np.array_equal(ph["tri"], ptri)),
                      "primary_uv_bits_equal_on_hits": bool(np.array_equal(ph["uv"].view(np.uint32)[ptri >= 0], puv.view(np.uint32)[ptri >= 0])),
                      "reference_glsl_closest_seconds": s_closest, "reference_glsl_any_seconds": s_any, "cpu_threads": cpu.hardware_threads(),
                      "reference_glsl_Mrays_per_s": rays.shape[0] / s_closest / 1e6}))

if __name__ == "__main__":
    main()
