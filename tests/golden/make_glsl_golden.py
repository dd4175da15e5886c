#!/usr/bin/env python3
"""Regenerates tests/golden/glsl_city12.npz: answers computed by the REFERENCE'S OWN SHADERS (shaders/traversal.glsl,
primaryray.glsl, pathtracer.glsl run on the CPU through oracle/_ref/libadypt_glsl.so, see oracle/glsl_transpile.py)
with the reference's own Sobol generator (oracle/_ref/libadypt_ref.so), on the reference-built city12 scene. The GPU
tests compare the CUDA kernels with these directly, so the chain reference shader -> fixture -> CUDA needs no oracle.

Run in the build container (needs /root/reference):   python tests/golden/make_glsl_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden  # noqa: E402
from oracle import glsl_ref, ref  # noqa: E402

W_, H_ = 96, 64
CFG_A = dict(max_bounce=5, subpixel=8, tmp_lifetime=16, ray_tmin=1e-4, clamp=4.0, sun=(1.0, 0.9, 0.8))
CFG_B = dict(max_bounce=3, subpixel=2, tmp_lifetime=5, ray_tmin=1e-4, clamp=2.0, sun=(5.0, 4.0, 3.0))
SPP_A, SPP_B = 40, 23


def main():
    g = load_golden("city12")
    cam = g.extra["cam"]
    m = ref.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), W_, H_)
    rays = g.extra["rays"]
    tri, uv = glsl_ref.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
    occ = glsl_ref.trace_any(g.nodes, g.woop, rays)
    out = dict(size=np.array([W_, H_], dtype=np.int32), glsl_tri=tri, glsl_uv=uv, glsl_any=occ)
    for vt in (0, 1, 2, 4, 5):
        out[f"view{vt}"] = glsl_ref.primary_view(g, cam[:3], 1e-4, m["inv_proj"], m["inv_view"], W_, H_, vt)
    bias = np.random.default_rng(20261017).integers(0, 256, size=(H_ * W_, 2), dtype=np.uint8)
    out["bias"] = bias
    for tag, cfg, spp in (("a", CFG_A, SPP_A), ("b", CFG_B, SPP_B)):
        sob = ref.sobol_sequence(2 * cfg["max_bounce"], spp)
        img, tmp = glsl_ref.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, sob, 0, spp)
        out["pt_" + tag] = img
        out["sobol_" + tag] = sob
        out["cfg_" + tag] = np.array([cfg["max_bounce"], cfg["subpixel"], cfg["tmp_lifetime"], cfg["ray_tmin"], cfg["clamp"], *cfg["sun"], spp], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "glsl_city12.npz"), **out)
    print("wrote glsl_city12.npz", {k: v.shape for k, v in out.items()}, os.path.getsize(os.path.join(HERE, "glsl_city12.npz")), "bytes")


if __name__ == "__main__":
    main()
