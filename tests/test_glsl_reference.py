"""The oracle against the REFERENCE'S OWN SHADER TEXT. oracle/_ref/libadypt_glsl.so is shaders/traversal.glsl,
primaryray.glsl and pathtracer.glsl compiled for the CPU from where they lie (oracle/glsl_transpile.py: syntax
rewrites + the three FP-policy rewrites of DESIGN.md §3), driven the way OglPathTracer::Trace drives them and fed by
the reference's own Sobol generator. Everything must agree BIT FOR BIT: hit ids, uv, any-hit bits, all viewer
images, path-traced images (every illum branch, the primary-hit cache, sub-pixel strata, textures).

Needs the prebuilt library (built where /root/reference exists; it travels to the GPU box inside oracle/_ref/). The
committed fixture tests/golden/glsl_city12.npz carries the same answers for the GPU tests (test_gpu_glsl_golden.py).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

glsl = pytest.importorskip("oracle.glsl_ref")
if not glsl.available():
    pytest.skip("oracle/_ref/libadypt_glsl.so not built (needs /root/reference)", allow_module_level=True)

W_, H_ = 96, 64
CFGS = [dict(max_bounce=5, subpixel=8, tmp_lifetime=16, ray_tmin=1e-4, clamp=4.0, sun=(1.0, 0.9, 0.8)),
        dict(max_bounce=3, subpixel=2, tmp_lifetime=5, ray_tmin=1e-4, clamp=2.0, sun=(5.0, 4.0, 3.0)),
        dict(max_bounce=1, subpixel=4, tmp_lifetime=1, ray_tmin=1e-4, clamp=100.0, sun=(1.0, 1.0, 1.0)),
        dict(max_bounce=8, subpixel=1, tmp_lifetime=3, ray_tmin=1e-3, clamp=0.5, sun=(0.0, 0.0, 0.0))]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def sobol(refmod, cpu, dim, n):
    """The reference's Sobol::Next vectors when its library is here, else the (reference-pinned) oracle's."""
    return refmod.sobol_sequence(dim, n) if refmod is not None else cpu.sobol_sequence(dim, n)


@pytest.fixture(scope="module")
def ref_or_none():
    from oracle import ref
    return ref if ref.available() else None


@pytest.mark.parametrize("name", ["tiny_two_triangles", "tiny_shared_edge", "tiny_strip", "tiny_deep", "city12", "c1_sample"])
def test_traversal_equals_reference_shader(cpu, c1, name):
    if name == "c1_sample":  # 4096 of C1's primary rays on the 65 536-triangle lattice
        s = c1[1]
        nodes, ti, woop, rays = s.nodes, s.tri_indices, s.woop, np.load(os.path.join(GOLDEN, "c1_sample.npz"))["rays"]
    else:
        g = load_golden(name)
        nodes, ti, woop, rays = g.nodes, g.tri_indices, g.woop, g.extra["rays"]
    o = cpu.trace_closest(nodes, ti, woop, rays)
    tri, uv = glsl.trace_closest(nodes, ti, woop, rays)
    hit = tri >= 0
    assert hit.any() and np.array_equal(o["tri"], tri)
    assert np.array_equal(bits(o["uv"])[hit], bits(uv)[hit])
    assert np.array_equal(glsl.trace_any(nodes, woop, rays), cpu.trace_any(nodes, woop, rays)["occluded"])


@pytest.mark.parametrize("seed", [1, 2])
def test_traversal_equals_reference_shader_on_degenerate_soups(cpu, seed):
    """Degenerate / duplicate / coplanar / huge triangles, rays starting on the geometry, axis-parallel rays and
    exactly-zero direction components (the shader clamps them to 2^-64)."""
    from adypt_b200 import host, workloads as W
    rng = np.random.default_rng(seed)
    n = 2000
    tri = rng.uniform(-4, 4, size=(n, 1, 3)) + rng.normal(0, 1, size=(n, 3, 3)) * rng.choice([0.01, 0.2, 1.5], size=(n, 1, 1))
    tri[:40, 2] = tri[:40, 1]
    tri[40:80, 1] = (tri[40:80, 0] + tri[40:80, 2]) * 0.5
    tri[100:200] = tri[200:300]
    tri[300:400, :, 2] = 0.5
    tri[400:410] *= 50.0
    pos = (np.round(tri * 64) / 64).astype(np.float32)
    hs = host.HostScene.from_triangles(pos, np.zeros(n, dtype=np.int32), host.materials_array(W.tiny_scene("strip").materials)).build_bvh()
    woop = cpu.build_woop(hs.tris, hs.tri_indices)
    rays = W.random_rays(30000, hs.aabb[:3] - 1, hs.aabb[3:] + 1, seed=seed)
    on = W.random_rays(10000, [-1, -1, -1], [1, 1, 1], seed=seed + 10)
    k = rng.integers(0, n, size=10000)
    on[:, :3] = (pos[k] * rng.dirichlet([1, 1, 1], size=10000).astype(np.float32)[:, :, None]).sum(axis=1)
    axis = W.random_rays(3000, hs.aabb[:3], hs.aabb[3:], seed=seed + 20)
    axis[:, 4:7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=3000)] * rng.choice([-1.0, 1.0], size=(3000, 1)).astype(np.float32)
    axis[2000:, 4:7] = np.where(axis[2000:, 4:7] == 0, np.float32(-0.0), axis[2000:, 4:7])
    rays = np.concatenate([rays, on, axis])
    o = cpu.trace_closest(hs.nodes, hs.tri_indices, woop, rays)
    tri_g, uv_g = glsl.trace_closest(hs.nodes, hs.tri_indices, woop, rays)
    hit = tri_g >= 0
    assert 0.05 < hit.mean() < 0.99
    assert np.array_equal(o["tri"], tri_g) and np.array_equal(bits(o["uv"])[hit], bits(uv_g)[hit])
    assert np.array_equal(glsl.trace_any(hs.nodes, woop, rays), cpu.trace_any(hs.nodes, woop, rays)["occluded"])


def test_viewer_images_equal_reference_shader(cpu):
    g = load_golden("city12")
    cam = g.extra["cam"]
    m = cpu.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), W_, H_)
    for vt in (0, 1, 2, 4, 5):
        a = cpu.primary_view(g, cam[:3], 1e-4, m["inv_proj"], m["inv_view"], W_, H_, vt)
        b = glsl.primary_view(g, cam[:3], 1e-4, m["inv_proj"], m["inv_view"], W_, H_, vt)
        assert np.array_equal(bits(a), bits(b)), vt
        assert a[:, :3].any()


@pytest.mark.parametrize("ci", range(len(CFGS)))
def test_path_traced_images_equal_reference_shader(cpu, ref_or_none, ci):
    g = load_golden("city12")
    illum = g.mats[:, 48:52].copy().view(np.int32).ravel()
    assert {1, 2, 3, 7} <= set(illum.tolist())  # every branch of pathtracer.glsl:144-201: diffuse, glossy, mirror, glass
    cfg = CFGS[ci]
    cam = g.extra["cam"]
    m = cpu.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), W_, H_)
    bias = np.random.default_rng(5 + ci).integers(0, 256, size=(H_ * W_, 2), dtype=np.uint8)
    spp = 37
    sob = sobol(ref_or_none, cpu, 2 * cfg["max_bounce"], spp)
    a, pa, _ = cpu.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, 0, spp)
    b, pb = glsl.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, sob, 0, spp)
    assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(bits(pa[:, [0, 2, 3]]), bits(pb[:, [0, 2, 3]]))  # uPrimaryTmpImg: id bits, u, v (.y is never written)
    assert cfg["sun"] == (0.0, 0.0, 0.0) or a[:, :3].mean() > 0.01
    # continuing an accumulation (frames 37..52 on top of the first 37) is the same on both sides
    a2, _, _ = cpu.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, spp, 16, out_rgba=a.copy(), primary_tmp=pa.copy())
    sob2 = sobol(ref_or_none, cpu, 2 * cfg["max_bounce"], spp + 16)
    b2, _ = glsl.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, sob2, spp, 16, out_rgba=b.copy(), primary_tmp=pb.copy())
    assert np.array_equal(bits(a2), bits(b2)) and not np.array_equal(bits(a2), bits(a))


def test_other_illum_values_equal_reference_shader(cpu, ref_or_none):
    """The remaining case labels (4, 5, 6) and values outside the switch (0, 9: the path goes straight on)."""
    g = load_golden("city12")
    mats = g.mats.copy()
    mats.view(np.int32).reshape(-1, 16)[:, 12] = [0, 4, 5, 6, 9, 1, 2, 6]
    g.mats = mats
    cfg = CFGS[0]
    cam = g.extra["cam"]
    m = cpu.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), W_, H_)
    bias = np.random.default_rng(77).integers(0, 256, size=(H_ * W_, 2), dtype=np.uint8)
    sob = sobol(ref_or_none, cpu, 2 * cfg["max_bounce"], 20)
    a, _, _ = cpu.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, 0, 20)
    b, _ = glsl.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, sob, 0, 20)
    assert np.array_equal(bits(a), bits(b)) and a[:, :3].mean() > 0.01


def test_textured_images_equal_reference_shader(cpu, ref_or_none, tmp_path):
    """texture(uTextures[m_dtex], texcoords).rgb with GL_REPEAT / GL_LINEAR, through both shaders."""
    pytest.importorskip("PIL.Image")
    from test_textures import textured_scene
    from adypt_b200 import workloads as W
    hs, textures = textured_scene(tmp_path)
    hs.woop = cpu.build_woop(hs.tris, hs.tri_indices)
    cam = W.city_camera(6)
    m = cpu.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], W_, H_)
    a = cpu.primary_view(hs, cam["position"], 1e-4, m["inv_proj"], m["inv_view"], W_, H_, 0, textures=textures)
    b = glsl.primary_view(hs, cam["position"], 1e-4, m["inv_proj"], m["inv_view"], W_, H_, 0, textures=textures)
    assert np.array_equal(bits(a), bits(b))
    assert not np.array_equal(a, cpu.primary_view(hs, cam["position"], 1e-4, m["inv_proj"], m["inv_view"], W_, H_, 0))
    cfg = CFGS[0]
    bias = np.random.default_rng(9).integers(0, 256, size=(H_ * W_, 2), dtype=np.uint8)
    sob = sobol(ref_or_none, cpu, 2 * cfg["max_bounce"], 24)
    pa, _, _ = cpu.pt_render(hs, cam["position"], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, 0, 24, textures=textures)
    pb, _ = glsl.pt_render(hs, cam["position"], m["inv_proj"], m["inv_view"], W_, H_, cfg, bias, sob, 0, 24, textures=textures)
    assert np.array_equal(bits(pa), bits(pb))


def test_committed_fixture_is_current(cpu, ref_or_none):
    """tests/golden/glsl_city12.npz (what the GPU tests compare with) still equals what the shaders produce."""
    z = np.load(os.path.join(GOLDEN, "glsl_city12.npz"))
    g = load_golden("city12")
    tri, uv = glsl.trace_closest(g.nodes, g.tri_indices, g.woop, g.extra["rays"])
    assert np.array_equal(tri, z["glsl_tri"]) and np.array_equal(bits(uv), bits(z["glsl_uv"]))
    cam = g.extra["cam"]
    m = cpu.camera_matrices(float(cam[5]), float(cam[3]), float(cam[4]), W_, H_)
    c = z["cfg_b"]
    cfg = dict(max_bounce=int(c[0]), subpixel=int(c[1]), tmp_lifetime=int(c[2]), ray_tmin=float(np.float32(c[3])), clamp=float(c[4]), sun=tuple(c[5:8]))
    img, _ = glsl.pt_render(g, cam[:3], m["inv_proj"], m["inv_view"], W_, H_, cfg, z["bias"], z["sobol_b"], 0, int(c[8]))
    assert np.array_equal(bits(img), bits(z["pt_b"]))
    assert np.array_equal(bits(cpu.sobol_sequence(6, int(c[8]))), bits(z["sobol_b"]))
