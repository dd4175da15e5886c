"""Full-size parity checks on the GPU box (BASELINE.json configs[2] and configs[3]); they replace the hand-run scripts under
tests/fullsize/ as far as the driver's `pytest -m gpu` is concerned.

  C3: the CUDA wavefront's 1920x1080 image after the first 16 samples per pixel must equal, bit for bit, the image the
      REFERENCE'S OWN pathtracer.glsl produced on the CPU -- held here only as a committed digest
      (tests/golden/hashes.json "c3_1080p_16spp", made by tests/golden/make_c3_digest.py where /root/reference exists):
      reference shader -> digest -> CUDA, no oracle in the chain.
  C4: ~10M triangles (622 MB of nodes + Woop rows, 5x the L2): ids / t / uv bit-identical to the oracle on a 250k-ray slice of the
      8M incoherent rays, any-hit == (closest id != -1) on all of them, shadow rays identical on a slice.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, fnv1a

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_c3_1080p_16spp_equals_reference_shader_digest(A):
    from adypt_b200 import host, workloads as W
    gold = json.load(open(os.path.join(GOLDEN, "hashes.json")))["c3_1080p_16spp"]
    w, h, spp = gold["width"], gold["height"], gold["spp"]
    mesh = W.city(183, 1, mixed_materials=True)
    scene = host.build_scene(mesh).upload(0)
    tr = A.Tracer(scene, A.PTConfig.make(sun=(1.0, 1.0, 1.0)), w, h, bias_seed=gold["bias_seed"])
    cam = W.city_camera(183)
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    tr.sample(spp)
    img = tr.read(4).reshape(-1, 4)
    assert abs(float(img[:, :3].astype(np.float64).mean()) - gold["mean_rgb"]) < 1e-9, "mean radiance differs from the reference shader's image"
    assert fnv1a(img) == gold["rgba_fnv"], "1920x1080 x 16 spp image differs from the reference's pathtracer.glsl image (digest)"
    # per-frame dispatch (one Trace(true) per sample, as the viewer's main loop does) gives the same image
    tr.primary(A.VIEW_DIFFUSE)
    for _ in range(spp):
        tr.sample(1)
    assert fnv1a(tr.read(4).reshape(-1, 4)) == gold["rgba_fnv"]
    tr.close()
    scene.close()


def test_c4_ten_million_triangles(A, cpu):
    import torch
    from adypt_b200 import host, workloads as W
    cells = 577
    mesh = W.city(cells, 1)
    hs = host.build_scene(mesh)  # own builder, 16 threads: ~17 s
    assert mesh.n_tris > 9_900_000
    scene = hs.upload(0)
    woop = cpu.build_woop(hs.tris, hs.tri_indices)
    assert np.array_equal(bits(scene.read_woop()), bits(woop))  # 10.4M Woop rows built on the GPU == glm::inverse order on the CPU
    tr = A.Tracer(scene, A.PTConfig.make(), 1000, 1000, bias_seed=7)
    cam = W.city_camera(cells)
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    prim = tr.primary_rays()
    ph = scene.trace_closest(prim)
    tr.close()
    rays = W.bounce_rays(mesh.positions(), prim, ph["tri"], ph["uv"], per_hit=8, seed=42)
    n = rays.shape[0]
    assert n > 4_000_000
    d_rays = torch.from_numpy(rays).cuda()
    d_tri = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    d_uv = torch.empty((n, 2), dtype=torch.float32, device="cuda")
    d_any = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    scene.trace_closest(d_rays, d_tri, d_t, d_uv, stream=st)
    scene.trace_any(d_rays, d_any, stream=st)
    torch.cuda.synchronize()
    tri, t, uv = d_tri.cpu().numpy(), d_t.cpu().numpy(), d_uv.cpu().numpy()
    sl = slice(2_000_000, 2_250_000)
    o = cpu.trace_closest(hs.nodes, hs.tri_indices, woop, rays[sl])
    assert np.array_equal(tri[sl], o["tri"])
    assert np.array_equal(bits(t[sl]), bits(o["t"]))
    assert np.array_equal(bits(uv[sl]), bits(o["uv"]))
    assert np.array_equal(d_any.cpu().numpy() != 0, tri >= 0), "any-hit and closest-hit disagree on which rays hit"
    st_gpu = scene.trace_stats(d_rays[: 8 * 250_000 * 0 + 2_250_000 * 8].view(-1)[2_000_000 * 8: 2_250_000 * 8].contiguous())
    assert st_gpu["nodes"] == o["counters"]["nodes"] and st_gpu["tris"] == o["counters"]["tris"]  # same work, ray by ray
    assert st_gpu["max_stack"] == o["counters"]["max_stack"] <= 8 + 56
    shadow = W.shadow_rays(mesh.positions(), tri, uv)[:250_000]
    assert np.array_equal(scene.trace_any(shadow), cpu.trace_any(hs.nodes, woop, shadow)["occluded"])
    scene.close()
