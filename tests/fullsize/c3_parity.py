#!/usr/bin/env python3
"""BASELINE.json configs[2] ("C3") at FULL size against the CPU oracle AND against the reference's own
pathtracer.glsl run on the CPU (oracle/_ref/libadypt_glsl.so, when built): 1920x1080, maxBounce 5, mixed-material
1M-triangle city, N spp (default 64). Prints the RMSE (expected: 0, the images are bit-identical) as JSON.
Test infrastructure: uses oracle/ as the checker. ~2-3 minutes of CPU time on the GPU box."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import adypt_b200 as A
from adypt_b200 import host, workloads as W
from oracle import cpu

def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    w, h = 1920, 1080
    mesh = W.city(183, 1, mixed_materials=True)
    hs = host.build_scene(mesh)
    sc = hs.upload(0)
    tr = A.Tracer(sc, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), w, h, bias_seed=7)
    cam = W.city_camera(183)
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    t0 = time.perf_counter(); tr.sample(spp); tr.sync(); gpu_s = time.perf_counter() - t0
    img = tr.read(4).reshape(-1, 4)
    hs.woop = cpu.build_woop(hs.tris, hs.tri_indices)
    m = cpu.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], w, h)
    cfg = dict(max_bounce=5, subpixel=8, tmp_lifetime=16, ray_tmin=1e-4, clamp=4.0, sun=(1.0, 1.0, 1.0))
    t0 = time.perf_counter()
    ref, _, cnt = cpu.pt_render(hs, cam["position"], m["inv_proj"], m["inv_view"], w, h, cfg, tr.get_bias(), 0, spp)
    cpu_s = time.perf_counter() - t0
    d = img[:, :3].astype(np.float64) - ref[:, :3].astype(np.float64)
    glsl_out = {}
    from oracle import glsl_ref
    if glsl_ref.available():
        sob = cpu.sobol_sequence(2 * cfg["max_bounce"], spp)  # == the reference's Sobol::Next (tests/test_oracle_pins.py)
        t0 = time.perf_counter()
        gimg, _ = glsl_ref.pt_render(hs, cam["position"], m["inv_proj"], m["inv_view"], w, h, cfg, tr.get_bias(), sob, 0, spp)
        glsl_out = {"reference_glsl_seconds": time.perf_counter() - t0,
                    "gpu_equals_reference_glsl": bool(np.array_equal(img.view(np.uint32), gimg.view(np.uint32))),
                    "oracle_equals_reference_glsl": bool(np.array_equal(ref.view(np.uint32), gimg.view(np.uint32))),
                    "rmse_vs_reference_glsl": float(np.sqrt(((img[:, :3].astype(np.float64) - gimg[:, :3]) ** 2).mean()))}
    print(json.dumps({**glsl_out, "config": f"C3: {w}x{h}, {spp} spp, maxBounce 5, mixed 1M-tri city", "rmse": float(np.sqrt((d ** 2).mean())),
                      "max_abs_diff": float(np.abs(d).max()), "bit_identical_pixels": float((img.view(np.uint32) == ref.view(np.uint32)).all(axis=1).mean()),
                      "bit_identical_image": bool(np.array_equal(img.view(np.uint32), ref.view(np.uint32))), "segments_gpu": tr.stats()["segments"],
                      "segments_oracle": cnt["segments"], "gpu_seconds": gpu_s, "oracle_seconds": cpu_s, "oracle_threads": cpu.hardware_threads(),
                      "mean_radiance": float(img[:, :3].mean())}))

if __name__ == "__main__":
    main()
