"""The product's from-scratch host stages against the reference's own C++ (oracle/_ref) and the
committed golden arrays: Triangle records, materials, SBVH nodes, 80-byte CWBVH nodes and leaf-order indices
must be BYTE-IDENTICAL; .bvh cache files must interoperate in both directions."""
import json
import os

import numpy as np
import pytest

from adypt_b200 import host, workloads as W
from conftest import CACHE, GOLDEN, fnv1a, load_golden


def same_as_reference(hs, b):
    assert np.array_equal(hs.tris, b.tris), "Triangle[] differs"
    assert np.array_equal(hs.mats, b.mats), "GPUMaterial[] differs"
    assert np.array_equal(hs.aabb, b.aabb)
    if b.sbvh_nodes is not None:
        assert np.array_equal(hs.binary_nodes, b.sbvh_nodes), "SBVH nodes differ"
    assert np.array_equal(hs.tri_indices, b.tri_indices), "leaf-order triangle indices differ"
    assert np.array_equal(hs.nodes, b.nodes), "CWBVH nodes differ"


@pytest.mark.parametrize("kind", ["two_triangles", "shared_edge", "strip", "deep"])
def test_tiny_scenes_byte_identical(refmod, kind):
    mesh = W.tiny_scene(kind)
    b = refmod.build(mesh.write_obj(CACHE), cache=False)
    same_as_reference(host.build_scene(mesh), b)


@pytest.mark.parametrize("cells,seed,mixed", [(12, 3, True), (24, 1, True), (31, 5, False)])
def test_city_with_spatial_splits_byte_identical(refmod, cells, seed, mixed):
    mesh = W.city(cells, seed, mixed_materials=mixed)
    b = refmod.build(mesh.write_obj(CACHE), cache=False)
    hs = host.build_scene(mesh)
    assert hs.tri_indices.size > hs.tris.shape[0]  # spatial splits duplicated references
    same_as_reference(hs, b)


@pytest.mark.parametrize("params", [(0, 0.3, 1.0), (48, 1.0, 1.0), (5, 0.1, 2.5)])
def test_builder_parameters_byte_identical(refmod, params):
    mesh = W.city(10, 7)
    b = refmod.build(mesh.write_obj(CACHE), *params, cache=False)
    same_as_reference(host.build_scene(mesh, host.BvhConfig.make(*params)), b)


def test_c1_byte_identical_and_matches_committed_digest(c1_ref):
    mesh, b = c1_ref
    hs = host.build_scene(mesh)
    same_as_reference(hs, b)
    h = json.load(open(os.path.join(GOLDEN, "hashes.json")))["c1"]
    assert fnv1a(hs.nodes) == h["nodes"] and fnv1a(hs.tri_indices) == h["tri_indices"] and fnv1a(hs.tris) == h["tris"]


def test_matches_golden_without_the_reference():
    """Runs anywhere (no oracle/_ref needed): city12's golden arrays came from the reference builder."""
    g = load_golden("city12")
    mesh = W.city(12, 3, mixed_materials=True, name="city12")
    hs = host.build_scene(mesh)
    assert np.array_equal(hs.nodes, g.nodes) and np.array_equal(hs.tri_indices, g.tri_indices)
    assert np.array_equal(hs.tris, g.tris) and np.array_equal(hs.mats, g.mats)


def test_obj_ingest_byte_identical(refmod, tmp_path):
    mesh = W.city(9, 11, mixed_materials=True)
    p = mesh.write_obj(str(tmp_path))
    b = refmod.build(p, cache=False)
    hs = host.HostScene.from_obj(p).build_bvh()
    same_as_reference(hs, b)


def test_obj_ingest_normals_texcoords_polygons_negative_indices(refmod, tmp_path):
    """vn / vt (v flipped, Scene.cpp:68-72), quads and a concave pentagon (ear clipping), negative indices,
    exponent notation, a material-less face (id -1)... the reference would index materials[-1] when SHADING
    such a face, but loading is defined."""
    obj = tmp_path / "m.obj"
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 0.1 0.2 0.3\nKe 1 2 3\nKs 0.5 0.5 0.5\nNs 96.078431\nNi 1.45\nd 0.5\nillum 2\n\n"
                                    "newmtl b\nKd 1 1 1\nTr 0.25\nillum 7\nmap_Kd tex.png\n")
    obj.write_text("""mtllib m.mtl
v 0 0 0
v 1.5e0 0 0
v 1.5 1 0
v 0 1 0.125
v 3 0 -0.333333343
v 4 0 1E-3
v 4 2 0
v 3.5 0.5 0
v 3 2 0
vn 0 0 1
vn 0 1 0
vt 0.25 0.75
vt 1 0
vt 0.5 0.5
usemtl a
f 1/1/1 2/2/1 3/3/2
f 1//1 2//1 3//1 4//2
usemtl b
f 5 6 7 8 9
f -5/1 -4/2 -3/3
usemtl missing
f 1 2 3
""")
    b = refmod.build(str(obj), cache=False)
    hs = host.HostScene.from_obj(str(obj))
    assert hs.tris.shape[0] == b.tris.shape[0] == 8
    assert np.array_equal(hs.tris, b.tris)
    assert np.array_equal(hs.mats, b.mats)


def test_bvh_cache_file_interoperates(refmod, tmp_path):
    mesh = W.city(12, 3, mixed_materials=True, name="city12")
    obj = mesh.write_obj(CACHE)
    hs = host.build_scene(mesh)
    ours, theirs = str(tmp_path / "ours.bvh"), str(tmp_path / "theirs.bvh")
    hs.save_bvh(ours)
    l = refmod.lib()
    h = l.ref_scene_load_bvh(obj.encode(), ours.encode(), 48, 0.3, 1.0)  # WideBVH::LoadFromFile reads our file
    assert h
    assert l.ref_scene_save_bvh(h, theirs.encode(), 48, 0.3, 1.0) == 0
    l.ref_scene_destroy(h)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    other = host.HostScene.from_triangles(mesh.positions(), mesh.face_mat, host.materials_array(mesh.materials))
    assert other.load_bvh(theirs) and np.array_equal(other.nodes, hs.nodes) and np.array_equal(other.tri_indices, hs.tri_indices)
    # reused only when all three build parameters match (WideBVH.cpp:42-45); missing / foreign files are rejected
    assert not other.load_bvh(theirs, host.BvhConfig.make(47, 0.3, 1.0))
    assert not other.load_bvh(theirs, host.BvhConfig.make(48, 0.31, 1.0))
    assert not other.load_bvh(str(tmp_path / "nope.bvh"))
    (tmp_path / "junk.bvh").write_bytes(b"NOTABVH\0" + bytes(64))
    assert not other.load_bvh(str(tmp_path / "junk.bvh"))
    assert l.ref_scene_load_bvh(obj.encode(), str(tmp_path / "junk.bvh").encode(), 48, 0.3, 1.0) is None


def test_single_triangle_is_rejected_not_crashed():
    """The reference builder reads child -1 on a one-triangle scene; ours reports it."""
    from adypt_b200 import AdyptError
    hs = host.HostScene.from_triangles(np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float32), np.zeros(1, dtype=np.int32),
                                       host.materials_array(W.tiny_scene("strip").materials))
    with pytest.raises(AdyptError):
        hs.build_bvh()


def test_traversal_of_own_bvh_matches_brute_force(cpu):
    mesh = W.city(8, 21)
    hs = host.build_scene(mesh)
    woop = cpu.build_woop(hs.tris, hs.tri_indices)
    rays = W.random_rays(2000, hs.aabb[:3] - 1, hs.aabb[3:] + 1, seed=3)
    r = cpu.trace_closest(hs.nodes, hs.tri_indices, woop, rays)
    b = cpu.brute_closest(hs.tri_indices, woop, rays)
    assert np.array_equal(r["t"].view(np.uint32), b["t"].view(np.uint32))
    assert (r["tri"] != b["tri"]).mean() < 1e-3 and (r["tri"] >= 0).sum() > 100
