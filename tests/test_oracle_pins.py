"""Pins the ORACLE (oracle/oracle.cpp) against the reference's own C++: committed golden vectors generated
from oracle/_ref (tests/golden/make_golden.py), and -- when oracle/_ref is present -- the live library."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, fnv1a, load_golden


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_sobol_matches_reference_golden(cpu):
    seq = np.load(os.path.join(GOLDEN, "ref_sobol.npz"))["seq"]  # Sobol::Next, 64 dims x 300 calls
    for dim in (2, 10, 33, 64):
        got = cpu.sobol_sequence(dim, 300)
        assert np.array_equal(bits(got), bits(seq[:, :dim]))
    # closed form used for sample sharding == sequential generator
    for idx in (0, 1, 2, 15, 16, 17, 255, 256, 299):
        assert np.array_equal(bits(cpu.sobol_at(64, idx)), bits(seq[idx]))
    assert np.all(seq[0] == 0.5)  # first vector is all 0.5


def test_sobol_all_10005_dimensions_golden(cpu):
    """Every dimension the reference carries (Sobol.hpp:9: m_x[10005], Sobol.inl:4: kMatrices[10005][32] with 10 000 rows
    listed and five zero rows): a digest of Sobol::Next over 10 005 dims x 300 calls, and the last 105 dimensions value by
    value, both generated from the reference's own generator (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "ref_sobol.npz"))
    got = cpu.sobol_sequence(10005, 300)
    assert fnv1a(got) == int(z["all_dims_fnv"]), "Sobol vectors over all 10 005 dimensions differ from the reference's"
    assert np.array_equal(bits(got[:, 9900:]), bits(z["tail"]))
    assert np.all(got[:, 10000:] == 0.0) and not np.all(got[1:, 9999] == 0.0)  # the five rows the reference leaves zero
    for idx in (0, 1, 255, 256, 299):
        assert np.array_equal(bits(cpu.sobol_at(10005, idx)), bits(got[idx]))


def test_sobol_dimension_limit(cpu):
    with pytest.raises(ValueError):
        cpu.sobol_sequence(10006, 1)


def test_sobol_directions_live(cpu, refmod):
    # direction numbers regenerated from Joe-Kuo parameters reproduce the reference generator: all 10 005 dimensions, 2 100 calls
    a = refmod.sobol_sequence(10005, 2100)
    b = cpu.sobol_sequence(10005, 2100)
    assert np.array_equal(bits(a), bits(b))


def test_camera_matrices_match_reference_golden(cpu):
    z = np.load(os.path.join(GOLDEN, "ref_camera.npz"))
    for i, (fov, yaw, pitch, w, h) in enumerate(z["params"]):
        m = cpu.camera_matrices(float(fov), float(yaw), float(pitch), int(w), int(h))
        for k in ("proj", "view", "inv_proj", "inv_view"):
            assert np.array_equal(bits(m[k]), bits(z[k][i])), (i, k)


def test_mat4_inverse_matches_reference_golden(cpu):
    z = np.load(os.path.join(GOLDEN, "ref_inverse.npz"))
    for m, inv in zip(z["m"], z["inv"]):
        assert np.array_equal(bits(cpu.mat4_inverse(m)), bits(inv))


@pytest.mark.parametrize("name", ["tiny_two_triangles", "tiny_shared_edge", "tiny_strip", "tiny_deep", "city12"])
def test_woop_matches_reference_golden(cpu, name):
    g = load_golden(name)  # woop here was built by the reference's glm::inverse (oracle/ref_wrap.cpp)
    w = cpu.build_woop(g.tris, g.tri_indices)
    assert np.array_equal(bits(w), bits(g.woop))


def test_woop_c1_live(cpu, c1_ref):
    _, b = c1_ref
    assert np.array_equal(bits(cpu.build_woop(b.tris, b.tri_indices)), bits(b.woop))


def test_reference_arrays_hashes_c1(c1_ref):
    """The reference builder run here reproduces the committed digests of its node / index / Woop arrays."""
    _, b = c1_ref
    h = json.load(open(os.path.join(GOLDEN, "hashes.json")))["c1"]
    assert (b.n_tris, b.n_nodes, b.n_refs) == (h["n_tris"], h["n_nodes"], h["n_refs"])
    assert fnv1a(b.nodes) == h["nodes"]
    assert fnv1a(b.tri_indices) == h["tri_indices"]
    assert fnv1a(b.woop) == h["woop"]
    assert fnv1a(b.tris) == h["tris"]


def test_c1_primary_hits_hash(cpu, c1):
    """C1: 1M coherent primary rays -- ray buffer and oracle hit ids reproduce the committed digests."""
    from adypt_b200 import workloads as W
    _, b = c1
    h = json.load(open(os.path.join(GOLDEN, "hashes.json")))["c1"]
    cam = W.lattice_camera()
    m = cpu.camera_matrices(cam["fov"], cam["yaw"], cam["pitch"], 1000, 1000)
    rays = cpu.primary_rays(cam["position"], 1e-4, m["inv_proj"], m["inv_view"], 1000, 1000)
    assert fnv1a(rays) == h["primary_rays"]
    r = cpu.trace_closest(b.nodes, b.tri_indices, b.woop, rays)
    assert int((r["tri"] >= 0).sum()) == h["hit_count"]
    assert fnv1a(r["tri"]) == h["hit_tri"]
    assert r["counters"]["max_stack"] <= 12  # the reference's default stackSize is enough here


def test_struct_layouts():
    """sizeof(WideBVHNode)=80, Woop=48, Triangle=100, GPUMaterial=64, InstanceConfig::PT=40."""
    import ctypes as C
    import adypt_b200 as A
    assert C.sizeof(A.PTConfig) == 40
    g = load_golden("city12")
    assert g.nodes.shape[1] == 80 and g.woop.shape[1] * 4 == 48 and g.tris.shape[1] == 100 and g.mats.shape[1] == 64
    # meta encoding of reference-built nodes: empty 0 | leaf unary count + offset < 24 | inner 001 + (24 + widx)
    meta = g.nodes[:, 24:32]
    imask = g.nodes[:, 15]
    for n in range(g.nodes.shape[0]):
        inner = [int(m) for m in meta[n] if m and (m & 0x1F) >= 24]
        assert all((m >> 5) == 1 for m in inner)
        assert sorted((m & 0x1F) - 24 for m in inner) == list(range(len(inner)))  # compacted ordinals
        assert int(imask[n]) == (1 << len(inner)) - 1
        for m in meta[n]:
            if m and (m & 0x1F) < 24:
                assert (m >> 5) in (0b001, 0b011, 0b111)
