"""GPU parity tests proper: the CUDA traversal, called through the C-ABI (adypt_trace_closest / adypt_trace_any),
against the CPU oracle on the same seeded inputs. Bit-exact ids, t and uv."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, fnv1a, load_golden

pytestmark = pytest.mark.gpu

NAMES = ["tiny_two_triangles", "tiny_shared_edge", "tiny_strip", "tiny_deep", "city12"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_same_hits(g, o):
    assert np.array_equal(g["tri"], o["tri"])
    assert np.array_equal(bits(g["t"]), bits(o["t"]))
    assert np.array_equal(bits(g["uv"]), bits(o["uv"]))


@pytest.mark.parametrize("name", NAMES)
def test_golden_scenes_closest_and_any(A, cpu, name):
    g = load_golden(name)
    sc = A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
    rays = g.extra["rays"]
    got = sc.trace_closest(rays)
    assert_same_hits(got, cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays))
    # and against the committed brute-force answers (ids may differ only on exact ties in t)
    assert np.array_equal(bits(got["t"]), bits(g.extra["exp_t"]))
    assert (got["tri"] != g.extra["exp_tri"]).mean() <= 1e-3
    occ = sc.trace_any(rays)
    assert np.array_equal(occ, cpu.trace_any(g.nodes, g.woop, rays)["occluded"])
    assert np.array_equal(occ != 0, g.extra["exp_tri"] >= 0)


@pytest.mark.parametrize("name", NAMES)
def test_gpu_built_woop_is_bit_identical(A, name):
    """woop = NULL: rows built on the GPU == rows the reference built with glm::inverse (golden)."""
    g = load_golden(name)
    sc = A.Scene(g.nodes, g.tri_indices, None, g.tris, g.mats)
    assert np.array_equal(bits(sc.read_woop()), bits(g.woop))


def test_known_answers(A):
    g = load_golden("tiny_two_triangles")
    sc = A.Scene(g.nodes, g.tri_indices, g.woop)  # traversal-only scene: no triangles / materials
    rays = np.array([[0.25, 0.25, 1.0, 1e-4, 0.0, 0.0, -1.0, 0.0], [3.25, 0.25, 3.0, 1e-4, 0.0, 0.0, -2.0, 0.0],
                     [0.25, 0.25, 1.0, 1e-4, 0.0, 0.0, 1.0, 0.0], [0.25, 0.25, 1.0, 1.5, 0.0, 0.0, -1.0, 0.0],
                     [0.75, 0.75, 1.0, 1e-4, 0.0, 0.0, -1.0, 0.0]], dtype=np.float32)
    r = sc.trace_closest(rays)
    assert r["tri"].tolist() == [0, 1, -1, -1, -1]
    assert np.allclose(r["t"][:2], [1.0, 2.0], rtol=1e-6) and np.all(r["t"][2:] == np.float32(1e9))
    assert np.allclose(r["uv"][:2], [[0.5, 0.25], [0.5, 0.25]], atol=1e-6)
    assert np.all(r["uv"][2:] == 0)
    assert sc.trace_any(rays).tolist() == [1, 1, 0, 0, 0]


def test_degenerate_directions(A, cpu):
    g = load_golden("tiny_deep")
    sc = A.Scene(g.nodes, g.tri_indices, g.woop)
    rays = np.array([[0.3, 0.3, -5.0, 1e-4, 0.0, 0.0, 1.0, 0.0], [0.3, -5.0, 0.3, 1e-4, -0.0, 1.0, 0.0, 0.0],
                     [-5.0, 0.3, 0.3, 1e-4, 1.0, 1e-30, -1e-30, 0.0], [0.3, 0.3, 0.3, 1e-4, 0.0, 0.0, 0.0, 0.0],
                     [0.3, 0.3, 0.3, 1e-4, 1e-38, -1e-40, 3.0, 0.0]], dtype=np.float32)
    assert_same_hits(sc.trace_closest(rays), cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays))


def test_empty_and_ragged_batches(A, cpu):
    g = load_golden("city12")
    sc = A.Scene(g.nodes, g.tri_indices, g.woop)
    rays = g.extra["rays"]
    o = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
    for n in (0, 1, 31, 32, 33, 255, 257, 4097):
        r = sc.trace_closest(rays[:n])
        assert r["tri"].shape == (n,)
        assert np.array_equal(r["tri"], o["tri"][:n]) and np.array_equal(bits(r["t"]), bits(o["t"][:n]))
        assert sc.trace_any(rays[:n]).shape == (n,)
    r = sc.trace_closest(rays[:100], want_t=False, want_uv=False)  # optional outputs may be NULL
    assert r["t"] is None and r["uv"] is None and np.array_equal(r["tri"], o["tri"][:100])


def test_kernel_variants_agree_bit_for_bit(A, cpu):
    """Every code-generation variant of the traversal kernel (scene.cu trace_kernel_for) -- the scalar kernel on the reference's 80-byte
    nodes (19), packed FFMA2 evaluations (13-15), packed on 96-byte nodes fetched with 256-bit loads (0 = the product, 16-18), the older
    tuning variants -- returns the oracle's ids, t and uv bits and any-hit flags."""
    for name in ("city12", "tiny_deep"):
        g = load_golden(name)
        sc = A.Scene(g.nodes, g.tri_indices, g.woop)
        rays = g.extra["rays"]
        want = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
        want_any = cpu.trace_any(g.nodes, g.woop, rays)["occluded"]
        for variant in (0, 19, 13, 14, 15, 16, 17, 18, 20, 21, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11):
            sc.configure(0, 0, variant)
            assert_same_hits(sc.trace_closest(rays), want)
            assert np.array_equal(sc.trace_any(rays), want_any), (name, variant)


def test_host_array_call_equals_device_call_across_its_chunk_schedule(A):
    """The host-array call cuts the batch into chunks with short first and last ones (scene.cu trace_host); sizes around the point
    where that schedule switches on, and odd ones that put chunk boundaries at odd ray indices, against ONE device launch."""
    import torch
    g = load_golden("city12")
    sc = A.Scene(g.nodes, g.tri_indices, g.woop)
    base = g.extra["rays"]
    n_max = 2 * (1 << 21) + 12345
    rays = np.ascontiguousarray(np.tile(base, ((n_max + base.shape[0] - 1) // base.shape[0], 1))[:n_max])
    rays[:, 0:3] += (np.arange(n_max, dtype=np.float32)[:, None] % 7.0) * np.float32(1e-3)  # not just copies of the fixture's rays
    d_rays = torch.from_numpy(rays).cuda()
    d_tri = torch.empty(n_max, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n_max, dtype=torch.float32, device="cuda")
    d_uv = torch.empty((n_max, 2), dtype=torch.float32, device="cuda")
    sc.trace_closest(d_rays, d_tri, d_t, d_uv)
    d_occ = torch.empty(n_max, dtype=torch.uint8, device="cuda")
    sc.trace_any(d_rays, d_occ)
    torch.cuda.synchronize()
    ref = {"tri": d_tri.cpu().numpy(), "t": d_t.cpu().numpy(), "uv": d_uv.cpu().numpy()}
    occ = d_occ.cpu().numpy()
    for n in (1966079, 1966080, 1966081, (1 << 21) - 1, 3000001, n_max):
        r = sc.trace_closest(rays[:n])
        assert np.array_equal(r["tri"], ref["tri"][:n]), n
        assert np.array_equal(bits(r["t"]), bits(ref["t"][:n])) and np.array_equal(bits(r["uv"]), bits(ref["uv"][:n])), n
        assert np.array_equal(sc.trace_any(rays[:n]), occ[:n]), n


def test_scheduling_knobs_do_not_change_results(A, cpu):
    """Refill threshold / CTAs per SM only change scheduling; per-ray results are invariant, as is any
    permutation of the ray buffer."""
    g = load_golden("city12")
    sc = A.Scene(g.nodes, g.tri_indices, g.woop)
    rays = g.extra["rays"]
    base = sc.trace_closest(rays)
    for ctas, thr in [(1, 1), (2, 16), (0, 32), (3, 24)]:
        sc.configure(ctas, thr)
        assert_same_hits(sc.trace_closest(rays), base)
    sc.configure(0, 0)
    perm = np.random.default_rng(0).permutation(rays.shape[0])
    p = sc.trace_closest(rays[perm])
    assert np.array_equal(p["tri"], base["tri"][perm]) and np.array_equal(bits(p["t"]), bits(base["t"][perm]))


def test_device_pointer_alignment_is_checked(A):
    import torch
    g = load_golden("tiny_strip")
    sc = A.Scene(g.nodes, g.tri_indices, g.woop)
    buf = torch.zeros(8 * 64 + 1, dtype=torch.float32, device="cuda")
    tri = torch.empty(64, dtype=torch.int32, device="cuda")
    with pytest.raises(A.AdyptError):
        sc.trace_closest(buf[1:], tri, None, None)  # 4-byte offset: not 16-byte aligned


def test_scene_validation_rejects_out_of_range_indices(A):
    g = load_golden("tiny_strip")
    bad = g.nodes.copy()
    bad[0, 16:20] = np.frombuffer(np.uint32(10_000).tobytes(), dtype=np.uint8)  # child_base past the node array
    with pytest.raises(A.AdyptError):
        A.Scene(bad, g.tri_indices, g.woop)
    with pytest.raises(A.AdyptError):
        A.Scene(g.nodes, g.tri_indices + 1000, None, g.tris, g.mats)


def test_scene_validation_covers_imask(A):
    """A node whose imask does not cover its inner children (a corrupt / stale .bvh cache) is refused: the kernel addresses a
    child as child_base + popc(imask & lowmask(ordinal)) and would read past the node array (traversal.glsl:61-67)."""
    g = load_golden("tiny_deep")
    bad = g.nodes.copy()
    assert bad[0, 15] != 0, "fixture's root has inner children"
    bad[0, 15] = 0  # imask byte of m_head.w
    with pytest.raises(A.AdyptError):
        A.Scene(bad, g.tri_indices, g.woop)
    bad = g.nodes.copy()
    last = np.frombuffer(np.uint32(g.nodes.shape[0] - 1).tobytes(), dtype=np.uint8)
    bad[0, 16:20] = last  # child_base = last node, but imask names more than one child
    if bin(int(g.nodes[0, 15])).count("1") > 1:
        with pytest.raises(A.AdyptError):
            A.Scene(bad, g.tri_indices, g.woop)


def test_material_id_out_of_range_can_be_traced_but_not_shaded(A, cpu):
    """An OBJ face before the first usemtl gets material id -1 (Scene.cpp:52, tinyobjloader): the reference's GL buffer read
    shrugs that off, a CUDA read would fault and poison the context. Traversal works; creating a tracer is refused."""
    g = load_golden("tiny_strip")
    tris = g.tris.copy()
    tris.view(np.int32).reshape(tris.shape[0], 25)[0, 24] = -1
    sc = A.Scene(g.nodes, g.tri_indices, None, tris, g.mats)
    rays = g.extra["rays"]
    got = sc.trace_closest(rays)
    exp = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
    assert np.array_equal(got["tri"], exp["tri"])
    with pytest.raises(A.AdyptError, match="material id"):
        A.Tracer(sc, A.PTConfig.make(), 16, 16, bias_seed=1)
    tris.view(np.int32).reshape(tris.shape[0], 25)[0, 24] = g.mats.shape[0]
    sc2 = A.Scene(g.nodes, g.tri_indices, None, tris, g.mats)
    with pytest.raises(A.AdyptError, match="material id"):
        A.Tracer(sc2, A.PTConfig.make(), 16, 16, bias_seed=1)


def test_c1_primary_rays_bit_exact(A, cpu, c1):
    """Config 1: 65 536-triangle lattice, 1M coherent primary rays: ids, t, uv identical to the oracle,
    and ids reproduce the committed digest."""
    from adypt_b200 import workloads as W
    _, b = c1
    h = json.load(open(os.path.join(GOLDEN, "hashes.json")))["c1"]
    sc = A.Scene(b.nodes, b.tri_indices, None, b.tris, b.mats)
    assert np.array_equal(bits(sc.read_woop()), bits(b.woop))
    tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=1)
    cam = W.lattice_camera()
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    rays = tr.primary_rays()
    assert fnv1a(rays) == h["primary_rays"]  # generate kernel == Camera() of primaryray.glsl, bit for bit
    got = sc.trace_closest(rays)
    assert fnv1a(got["tri"]) == h["hit_tri"] and int((got["tri"] >= 0).sum()) == h["hit_count"]
    assert_same_hits(got, cpu.trace_closest(b.nodes, b.tri_indices, b.woop, rays))


def test_c2_incoherent_rays(A, cpu, c2):
    """Config 2: ~1M-triangle city. 8M incoherent bounce rays are traced on the GPU; a 1M-ray slice is checked
    against the oracle (>= 99.99 % equal ids required; we get and assert 100 %, t bit-exact), and the full
    8M set through size-independent properties: any-hit == (closest id != -1), host call == device call,
    determinism."""
    import torch
    from adypt_b200 import workloads as W
    mesh, b = c2
    sc = A.Scene(b.nodes, b.tri_indices, b.woop, b.tris, b.mats)
    tr = A.Tracer(sc, A.PTConfig.make(), 1000, 1000, bias_seed=1)
    cam = W.city_camera(183)
    tr.look(cam["position"], cam["yaw"], cam["pitch"], cam["fov"])
    prim = tr.primary_rays()
    ph = sc.trace_closest(prim)
    rays = W.bounce_rays(mesh.positions(), prim, ph["tri"], ph["uv"], per_hit=8, seed=42)
    n = rays.shape[0]
    assert n == 8_000_000
    d_rays = torch.from_numpy(rays).cuda()
    d_tri = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    d_uv = torch.empty((n, 2), dtype=torch.float32, device="cuda")
    d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    sc.trace_closest(d_rays, d_tri, d_t, d_uv, stream=st)
    sc.trace_any(d_rays, d_occ, stream=st)
    torch.cuda.synchronize()
    tri, t, occ = d_tri.cpu().numpy(), d_t.cpu().numpy(), d_occ.cpu().numpy()
    sl = slice(3_000_000, 4_000_000)
    o = cpu.trace_closest(b.nodes, b.tri_indices, b.woop, rays[sl])
    match = (tri[sl] == o["tri"]).mean()
    assert match >= 0.9999, match
    assert match == 1.0
    rel = np.abs(t[sl] - o["t"]) / np.maximum(np.abs(o["t"]), 1e-30)
    assert rel.max() <= 1e-5 and np.array_equal(bits(t[sl]), bits(o["t"]))
    assert np.array_equal(bits(d_uv.cpu().numpy()[sl]), bits(o["uv"]))
    # full-size properties
    assert np.array_equal(occ != 0, tri >= 0)
    assert 0.3 < (tri >= 0).mean() < 0.9
    again = torch.empty_like(d_tri)
    sc.trace_closest(d_rays, again, None, None, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(again, d_tri)
    host = sc.trace_closest(rays[:500_000])
    assert np.array_equal(host["tri"], tri[:500_000]) and np.array_equal(bits(host["t"]), bits(t[:500_000]))
    # a ray's result does not depend on its neighbours: all 8M rays in a random order give the same results, permuted (which
    # warp and lane a ray lands in, who refills when, is scheduling only)
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    p_rays = d_rays[perm].contiguous()
    p_tri, p_t, p_uv = torch.empty_like(d_tri), torch.empty_like(d_t), torch.empty_like(d_uv)
    sc.trace_closest(p_rays, p_tri, p_t, p_uv, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(p_tri, d_tri[perm]) and torch.equal(p_t.view(torch.int32), d_t[perm].view(torch.int32))
    assert torch.equal(p_uv.view(torch.int32), d_uv[perm].view(torch.int32))
    # the structure-of-arrays entry the wavefront uses (origins and directions as two arrays) is the same kernel with another stride
    for variant in (0, 19, 13, 9):  # product kernel; the scalar kernel on 80-byte nodes; packed on 80-byte nodes; hit-mask table variant
        sc.configure(0, 0, variant)
        sc.trace_closest(d_rays[:1_000_000], again[:1_000_000], None, None, stream=st)
        torch.cuda.synchronize()
        assert torch.equal(again[:1_000_000], d_tri[:1_000_000]), variant
    sc.configure(0, 0, 0)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_triangle_soups(A, cpu, seed):
    """Random triangle soups with the nasty cases mixed in -- degenerate (zero-area) triangles whose Woop rows are
    inf/NaN, coplanar duplicates (exact ties in t), slivers, huge and tiny triangles, axis-aligned quads sharing
    edges -- built by the product's own builder, traced on the GPU and by the oracle: ids, t, uv and any-hit bits
    must be identical for rays from everywhere, including rays that start on the geometry."""
    from adypt_b200 import host, workloads as W
    rng = np.random.default_rng(seed)
    n = 3000
    c = rng.uniform(-4, 4, size=(n, 1, 3))
    tri = c + rng.normal(0, 1, size=(n, 3, 3)) * rng.choice([0.01, 0.2, 1.5], size=(n, 1, 1))
    tri[:50, 2] = tri[:50, 1]                       # degenerate: two equal corners
    tri[50:100, 1] = (tri[50:100, 0] + tri[50:100, 2]) * 0.5  # degenerate: collinear
    tri[100:200] = tri[200:300]                     # exact duplicates
    tri[300:400, :, 2] = 0.5                        # coplanar in z = 0.5
    tri[400:420] *= 50.0                            # huge
    tri = np.round(tri * 64) / 64                   # snap to a grid: many exactly shared coordinates
    pos = tri.astype(np.float32)
    mats = host.materials_array(W.tiny_scene("strip").materials)
    hs = host.HostScene.from_triangles(pos, np.zeros(n, dtype=np.int32), mats).build_bvh()
    woop = cpu.build_woop(hs.tris, hs.tri_indices)
    sc = hs.upload(0)
    got_woop = sc.read_woop()
    # degenerate triangles give inf / NaN rows: same places on both sides (NaN payload bits are not specified by
    # IEEE 754 -- x86 produces 0xFFC00000, the GPU 0x7FFFFFFF -- and no comparison can observe them)
    nan_g, nan_o = np.isnan(got_woop), np.isnan(woop)
    assert np.array_equal(nan_g, nan_o) and nan_o.any()
    assert np.array_equal(got_woop.view(np.uint32)[~nan_o], woop.view(np.uint32)[~nan_o])
    rays = W.random_rays(60000, hs.aabb[:3] - 1, hs.aabb[3:] + 1, seed=seed)
    on = W.random_rays(20000, [-1, -1, -1], [1, 1, 1], seed=seed + 10)   # origins ON triangles (bounce-like rays)
    k = rng.integers(0, n, size=20000)
    bary = rng.dirichlet([1, 1, 1], size=20000).astype(np.float32)
    on[:, :3] = (pos[k] * bary[:, :, None]).sum(axis=1)
    axis = W.random_rays(3000, hs.aabb[:3], hs.aabb[3:], seed=seed + 20)  # exactly axis-parallel directions
    axis[:, 4:7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=3000)] * rng.choice([-1.0, 1.0], size=(3000, 1)).astype(np.float32)
    rays = np.concatenate([rays, on, axis])
    g = sc.trace_closest(rays)
    o = cpu.trace_closest(hs.nodes, hs.tri_indices, woop, rays)
    assert np.array_equal(g["tri"], o["tri"])
    assert np.array_equal(bits(g["t"]), bits(o["t"])) and np.array_equal(bits(g["uv"]), bits(o["uv"]))
    assert np.array_equal(sc.trace_any(rays), cpu.trace_any(hs.nodes, woop, rays)["occluded"])
    assert 0.05 < (g["tri"] >= 0).mean() < 0.99
    assert sc.trace_stats(rays)["max_stack"] == o["counters"]["max_stack"]
    assert sc.trace_stats(rays)["nodes"] == o["counters"]["nodes"] and sc.trace_stats(rays)["tris"] == o["counters"]["tris"]


def test_stack_deeper_than_the_shared_memory_part(A, cpu):
    """Clusters of triangles on a geometric progression of scales give a wide tree that is a long chain of nodes with
    several inner children each; rays along the chain defer a sibling group at every level, so the traversal stack grows
    past the 8 entries a lane keeps in shared memory (reference default stackSize: 12, unchecked) and uses the
    local-memory part. Results and counters must still equal the oracle's."""
    from adypt_b200 import host, workloads as W
    levels, per, ratio = 120, 64, 1.25
    rng = np.random.default_rng(3)
    tris = []
    for i in range(levels):
        s = ratio ** i
        c = np.array([3.0 * s, 0.0, 0.0])
        for _ in range(per):
            p = c + (rng.random(3) - 0.5) * s * 0.8
            tris.append([p, p + (rng.random(3) - 0.5) * s * 0.3, p + (rng.random(3) - 0.5) * s * 0.3])
    pos = np.array(tris, dtype=np.float32)
    mats = host.materials_array(W.tiny_scene("strip").materials)
    hs = host.HostScene.from_triangles(pos, np.zeros(len(pos), dtype=np.int32), mats).build_bvh()
    n = 4000
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0] = 3.0 * ratio ** (levels - 1) * 1.5
    rays[:, 1:3] = (rng.random((n, 2)) - 0.5) * 0.2
    rays[:, 3] = 1e-4
    target = np.zeros((n, 3))
    target[:, 0] = 3.0
    target[:, 1:] = (rng.random((n, 2)) - 0.5) * 0.5
    d = target - rays[:, :3]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    back = rays.copy()
    back[:, 0] = 0.0
    back[:, 4:7] = -d
    rays = np.concatenate([rays, back])
    woop = cpu.build_woop(hs.tris, hs.tri_indices)
    o = cpu.trace_closest(hs.nodes, hs.tri_indices, woop, rays)
    assert o["counters"]["max_stack"] > 12
    sc = hs.upload(0)
    assert_same_hits(sc.trace_closest(rays), o)
    assert np.array_equal(sc.trace_any(rays), cpu.trace_any(hs.nodes, woop, rays)["occluded"])
    st = sc.trace_stats(rays)
    assert st["max_stack"] == o["counters"]["max_stack"]
    assert st["nodes"] == o["counters"]["nodes"] and st["tris"] == o["counters"]["tris"]
