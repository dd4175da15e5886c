#!/usr/bin/env python3
"""Regenerates the committed fixtures in tests/golden/ from the REFERENCE's own C++ (oracle/_ref, compiled in
place from /root/reference) and, for traversal answers, from the brute-force Woop test in the oracle.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

Fixtures (all small):
  ref_sobol.npz      Sobol::Next output, 300 calls: dims 1-64, dims 9901-10005, digest of all 10005              (src/Util/Sobol.cpp:16-21)
  ref_camera.npz     Camera::GetView/GetProjection + glm::inverse for 6 cameras (src/Tracer/Camera.cpp:13-23)
  ref_inverse.npz    glm::inverse on 256 random mat4
  tiny_<kind>.npz    reference-built CWBVH arrays (nodes, tri_indices, woop, tris, mats) of hand-checkable
                     scenes + rays + brute-force answers (tri, t, uv)
  city12.npz         reference-built arrays of a 12x12-cell mixed-material city (~4k triangles) + 4096 rays
                     (primary + bounce) with brute-force answers, + camera for render tests
  c1_sample.npz      4096 of C1's 1M primary rays with brute-force answers
  hashes.json        FNV digests of the reference-built node / index / Woop arrays of C1 and C2's procedural
                     OBJ (so "uploaded unchanged" is checkable) and of the oracle's C1 hit ids
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from adypt_b200 import workloads as W  # noqa: E402
from conftest import CACHE, fnv1a  # noqa: E402
from oracle import cpu, ref  # noqa: E402


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


def scene_arrays(b):
    return dict(nodes=b.nodes, tri_indices=b.tri_indices, woop=b.woop, tris=b.tris, mats=b.mats)


def brute(b, rays):
    r = cpu.brute_closest(b.tri_indices, b.woop, rays)
    return dict(rays=rays, exp_tri=r["tri"], exp_t=r["t"], exp_uv=r["uv"])


def main():
    full = ref.sobol_sequence(10005, 300)  # every dimension of the reference's table (Sobol.hpp:9)
    save("ref_sobol", seq=full[:, :64].copy(), tail=full[:, 9900:].copy(), all_dims_fnv=np.uint64(fnv1a(full)))
    cams = [(45.0, 38.0, -24.0, 1000, 1000), (45.0, 7.0, -31.0, 1920, 1080), (45.0, 0.0, 0.0, 1280, 720),
            (33.3, 181.5, 89.0, 640, 480), (90.0, 359.0, -89.5, 3840, 2160), (60.0, 123.4, 12.3, 333, 777)]
    mats = [ref.camera_matrices(*c) for c in cams]
    save("ref_camera", params=np.array(cams, dtype=np.float64), proj=np.stack([m["proj"] for m in mats]),
         view=np.stack([m["view"] for m in mats]), inv_proj=np.stack([m["inv_proj"] for m in mats]),
         inv_view=np.stack([m["inv_view"] for m in mats]))
    rng = np.random.default_rng(1234)
    m_in = rng.standard_normal((256, 16)).astype(np.float32)
    m_in[:16] *= 1e-3
    m_in[16:32] *= 1e3
    save("ref_inverse", m=m_in, inv=np.stack([ref.mat4_inverse(m) for m in m_in]))

    for kind in ("two_triangles", "shared_edge", "strip", "deep"):
        mesh = W.tiny_scene(kind)
        b = ref.build(mesh.write_obj(CACHE), cache=False)
        lo, hi = b.aabb[:3] - 1.0, b.aabb[3:] + 1.0
        rays = W.random_rays(512, lo, hi, seed=11)
        # rays aimed at the geometry from outside, so most of them hit
        tgt = W.random_rays(512, b.aabb[:3], b.aabb[3:], seed=12)[:, :3]
        org = W.random_rays(512, lo - 3.0, hi + 3.0, seed=13)[:, :3]
        aimed = np.zeros((512, 8), dtype=np.float32)
        aimed[:, :3] = org
        aimed[:, 3] = 1e-4
        aimed[:, 4:7] = tgt - org
        rays = np.concatenate([rays, aimed])
        save("tiny_" + kind, **scene_arrays(b), **brute(b, rays))

    mesh = W.city(12, 3, mixed_materials=True, name="city12")
    b = ref.build(mesh.write_obj(CACHE), cache=False)
    cam = dict(position=(6.3, 5.0, 15.5), yaw=4.0, pitch=-28.0, fov=45.0)
    m = ref.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], 64, 64)
    prim = cpu.primary_rays(cam["position"], 1e-4, m["inv_proj"], m["inv_view"], 64, 64)
    hit = cpu.trace_closest(b.nodes, b.tri_indices, b.woop, prim)
    bounce = W.bounce_rays(mesh.positions(), prim, hit["tri"], hit["uv"], per_hit=2, seed=42)[:4096]
    rays = np.concatenate([prim, bounce])
    save("city12", **scene_arrays(b), **brute(b, rays), cam=np.array([*cam["position"], cam["yaw"], cam["pitch"], cam["fov"]], dtype=np.float32))

    hashes = {}
    mesh = W.sphere_lattice(5)
    b = ref.build(mesh.write_obj(CACHE))
    cam = W.lattice_camera()
    m = ref.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], 1000, 1000)
    prim = cpu.primary_rays(cam["position"], 1e-4, m["inv_proj"], m["inv_view"], 1000, 1000)
    hit = cpu.trace_closest(b.nodes, b.tri_indices, b.woop, prim)
    sel = np.random.default_rng(5).choice(prim.shape[0], 4096, replace=False)
    sel.sort()
    bf = brute(b, prim[sel])
    assert np.array_equal(bf["exp_tri"], hit["tri"][sel]), "oracle traversal disagrees with brute force on C1 sample"
    save("c1_sample", index=sel, **bf)
    hashes["c1"] = dict(n_tris=int(b.n_tris), n_nodes=int(b.n_nodes), n_refs=int(b.n_refs), nodes=fnv1a(b.nodes),
                        tri_indices=fnv1a(b.tri_indices), woop=fnv1a(b.woop), tris=fnv1a(b.tris),
                        primary_rays=fnv1a(prim), hit_tri=fnv1a(hit["tri"]), hit_count=int((hit["tri"] >= 0).sum()))
    mesh = W.city(183, 1)
    b = ref.build(mesh.write_obj(CACHE))
    hashes["c2"] = dict(n_tris=int(b.n_tris), n_nodes=int(b.n_nodes), n_refs=int(b.n_refs), nodes=fnv1a(b.nodes),
                        tri_indices=fnv1a(b.tri_indices), woop=fnv1a(b.woop), tris=fnv1a(b.tris))
    with open(os.path.join(HERE, "hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    print(json.dumps(hashes, indent=1))


if __name__ == "__main__":
    main()
