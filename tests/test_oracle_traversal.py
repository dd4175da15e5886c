"""The oracle's traversal (restating shaders/traversal.glsl) against an independent O(N) brute-force Woop test
over all leaf references, on committed fixtures, plus hand-checked known answers and edge cases."""
import numpy as np
import pytest

from conftest import load_golden

NAMES = ["tiny_two_triangles", "tiny_shared_edge", "tiny_strip", "tiny_deep", "city12", ]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def check_against_brute(r, g, allow_ties=True):
    exp_tri, exp_t, exp_uv = g.extra["exp_tri"], g.extra["exp_t"], g.extra["exp_uv"]
    # hit distance must agree bit for bit; the id may differ only on an exact tie in t (first-tested wins,
    # traversal.glsl:235, and traversal order != leaf order)
    assert np.array_equal(bits(r["t"]), bits(exp_t))
    diff = r["tri"] != exp_tri
    if not allow_ties:
        assert not diff.any()
    assert diff.mean() <= 1e-3
    same = ~diff
    assert np.array_equal(bits(r["uv"][same]), bits(exp_uv[same]))


@pytest.mark.parametrize("name", NAMES)
def test_closest_matches_brute_force(cpu, name):
    g = load_golden(name)
    r = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, g.extra["rays"])
    check_against_brute(r, g)
    assert (r["tri"] >= 0).sum() > 0


@pytest.mark.parametrize("name", NAMES)
def test_any_hit_consistent_with_closest(cpu, name):
    g = load_golden(name)
    a = cpu.trace_any(g.nodes, g.woop, g.extra["rays"])
    assert np.array_equal(a["occluded"] != 0, g.extra["exp_tri"] >= 0)


def test_c1_sample_matches_brute_force(cpu, c1):
    from conftest import GoldenScene
    import os
    from conftest import GOLDEN
    _, b = c1
    z = np.load(os.path.join(GOLDEN, "c1_sample.npz"))
    r = cpu.trace_closest(b.nodes, b.tri_indices, b.woop, z["rays"])
    assert np.array_equal(r["tri"], z["exp_tri"])
    assert np.array_equal(bits(r["t"]), bits(z["exp_t"]))
    assert np.array_equal(bits(r["uv"]), bits(z["exp_uv"]))


def test_known_answer_single_triangle(cpu):
    """Hand-checked: ray straight down onto triangle 0 = (0,0,0),(1,0,0),(0,1,0) at (0.25,0.25):
    the point is 0.5*p1 + 0.25*p2 + 0.25*p3, so (u,v) = (0.5,0.25), t = 1 (weights of FetchInfo,
    pathtracer.glsl:77-85)."""
    g = load_golden("tiny_two_triangles")
    rays = np.array([[0.25, 0.25, 1.0, 1e-4, 0.0, 0.0, -1.0, 0.0],      # hits tri 0
                     [3.25, 0.25, 3.0, 1e-4, 0.0, 0.0, -2.0, 0.0],      # hits tri 1 at t = 2 (dir gets normalised)
                     [0.25, 0.25, 1.0, 1e-4, 0.0, 0.0, 1.0, 0.0],       # points away
                     [0.25, 0.25, 1.0, 1.5, 0.0, 0.0, -1.0, 0.0],       # tmin beyond the hit
                     [0.75, 0.75, 1.0, 1e-4, 0.0, 0.0, -1.0, 0.0]],     # outside the triangle (u+v>1 side)
                    dtype=np.float32)
    r = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
    assert r["tri"].tolist() == [0, 1, -1, -1, -1]
    assert np.allclose(r["t"][:2], [1.0, 2.0], rtol=1e-6)
    assert np.allclose(r["uv"][0], [0.5, 0.25], atol=1e-6)
    assert np.allclose(r["uv"][1], [0.5, 0.25], atol=1e-6)
    assert np.all(r["t"][2:] == np.float32(1e9))  # miss keeps hit_t = 1e9 (traversal.glsl:28)
    a = cpu.trace_any(g.nodes, g.woop, rays)
    assert a["occluded"].tolist() == [1, 1, 0, 0, 0]


def test_shared_edge_first_tested_wins(cpu):
    """Two coplanar triangles sharing the diagonal; a ray through the shared edge must resolve to the SAME
    triangle in traversal and brute force (single leaf group => identical test order, strict '<')."""
    g = load_golden("tiny_shared_edge")
    rays = np.array([[0.5, 0.5, 1.0, 1e-4, 1e-9, 1e-9, -1.0, 0.0],
                     [0.25, 0.25, 2.0, 1e-4, 1e-9, -1e-9, -1.0, 0.0]], dtype=np.float32)
    r = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
    b = cpu.brute_closest(g.tri_indices, g.woop, rays)
    assert np.array_equal(r["tri"], b["tri"]) and (r["tri"] >= 0).all()
    assert np.array_equal(bits(r["t"]), bits(b["t"]))


def test_degenerate_directions_are_clamped(cpu):
    """Exact-zero / denormal / negative-zero direction components are clamped to +-2^-64
    (traversal.glsl:16-19); results stay finite and equal to brute force."""
    g = load_golden("tiny_deep")
    rays = np.array([[0.3, 0.3, -5.0, 1e-4, 0.0, 0.0, 1.0, 0.0],
                     [0.3, -5.0, 0.3, 1e-4, -0.0, 1.0, 0.0, 0.0],
                     [-5.0, 0.3, 0.3, 1e-4, 1.0, 1e-30, -1e-30, 0.0],
                     [0.3, 0.3, 0.3, 1e-4, 0.0, 0.0, 0.0, 0.0]], dtype=np.float32)
    r = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
    b = cpu.brute_closest(g.tri_indices, g.woop, rays)
    assert np.isfinite(r["t"]).all()
    assert np.array_equal(bits(r["t"]), bits(b["t"]))


def test_empty_and_ragged_batches(cpu):
    g = load_golden("tiny_strip")
    rays = g.extra["rays"]
    full = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays)
    for n in (0, 1, 31, 33, 257):
        r = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, rays[:n])
        assert r["tri"].shape == (n,)
        assert np.array_equal(r["tri"], full["tri"][:n])


def test_thread_count_does_not_change_results(cpu):
    g = load_golden("city12")
    a = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, g.extra["rays"], nthreads=1)
    b = cpu.trace_closest(g.nodes, g.tri_indices, g.woop, g.extra["rays"], nthreads=5)
    assert np.array_equal(a["tri"], b["tri"]) and np.array_equal(bits(a["t"]), bits(b["t"]))
    assert a["counters"]["nodes"] == b["counters"]["nodes"] and a["counters"]["tris"] == b["counters"]["tris"]
