#!/usr/bin/env python3
"""Digest of the reference's OWN pathtracer.glsl image for BASELINE.json configs[2] at full resolution: 1920x1080, the first 16
samples per pixel (one tmpLifetime block), maxBounce 5, mixed-material 1M-triangle city, bias seed 7. Everything on the
reference side: scene arrays from the reference's OBJ -> SBVH -> CWBVH pipeline (oracle/_ref/libadypt_ref.so), Sobol vectors
from its Sobol::Next, the image from its shaders compiled for the CPU (oracle/_ref/libadypt_glsl.so). Writes the FNV digest of
the RGBA float image into tests/golden/hashes.json ("c3_1080p_16spp"); tests/test_gpu_fullsize.py holds the CUDA render to it.
Run where /root/reference exists (about a minute on 8 cores):  python tests/golden/make_c3_digest.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402  (workload constants + reference_bias)
from conftest import fnv1a  # noqa: E402
from oracle import cpu, glsl_ref, ref  # noqa: E402


def main():
    spp = 16
    w, h = bench.C3["width"], bench.C3["height"]
    _, bvh = bench.reference_inputs(mixed=True)
    cam = bench.W.city_camera(bench.CELLS)
    m = ref.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], w, h)
    sob = ref.sobol_sequence(2 * bench.PT_CFG["max_bounce"], spp)
    bias = bench.reference_bias(w * h, seed=7)
    img, _ = glsl_ref.pt_render(bvh, cam["position"], m["inv_proj"], m["inv_view"], w, h, bench.PT_CFG, bias, sob, 0, spp)
    p = os.path.join(HERE, "hashes.json")
    hashes = json.load(open(p))
    hashes["c3_1080p_16spp"] = {"width": w, "height": h, "spp": spp, "bias_seed": 7, "rgba_fnv": fnv1a(img), "mean_rgb": float(img[:, :3].astype(np.float64).mean()),
                                "source": "reference pathtracer.glsl on the CPU (tests/golden/make_c3_digest.py)"}
    json.dump(hashes, open(p, "w"), indent=1)
    print(hashes["c3_1080p_16spp"])


if __name__ == "__main__":
    main()
