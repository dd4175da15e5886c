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


def _fuzz_obj(rnd, d):
    """One random OBJ + MTL pair in the spellings real exporters and hand-edited files produce -- and some they
    should not: odd number formats, malformed numbers, tabs, doubled and trailing blanks, LF / CRLF / CR line
    ends, negative indices, polygons, all face corner forms, unknown / empty / repeated material names, several
    mtllib names, keys before the first newmtl, texture options."""
    def num():
        v = rnd.uniform(-5, 5)
        if rnd.random() < 0.07:
            return rnd.choice(["1e", "e5", "--1", "1.2.3", "nan", "inf", "0x1p3", "1,5", "+", "-", ".", "1e+", "1.e2", "1.5e-3x", "3f", "1d2"])
        return rnd.choice([f"{v:.6f}", f"{v:.3e}", f"{v:.2E}", f"+{abs(v):.3f}", f"{int(v)}", f"{int(v)}.", f".{rnd.randrange(1000):03d}",
                           f"-.{rnd.randrange(100):02d}", f"{v:.9g}", f"{int(v)}e1", "0", "-0"])
    ws = lambda: rnd.choice([" ", "  ", "\t", " \t "])
    eol = lambda: rnd.choice(["\n", "\r\n", " \n", "\t\r\n", "\r", "\n\n"])
    gap = lambda: rnd.choice([" ", " ", " ", "  ", "\t"])
    nv, nvt, nvn = rnd.randrange(5, 12), rnd.randrange(0, 5), rnd.randrange(0, 5)
    names = ["m0", "m1", "mat_2"] + (["m0"] if rnd.random() < 0.1 else [])
    mtl = ["Kd 1 2 3"] if rnd.random() < 0.15 else []
    for nm in names:
        mtl.append("newmtl" + gap() + nm + rnd.choice(["", "", " ", "\t"]))
        if rnd.random() < 0.2:
            mtl.append(rnd.choice(["map_Kd tex.png", "map_Kd -s 1 1 1 tex.png", "map_Kd -o 0.5 0.5 tex.png", "map_Kd -clamp on -bm 2 my tex.png",
                                   "map_Kd  -blendu off tex.png ", "map_Kd -mm 0 1", "map_Kd -type sphere tex.png"]))
        keys = [("Kd", 3), ("Ke", 3), ("Ks", 3), ("Ns", 1), ("Ni", 1), ("d", 1), ("illum", 0), ("Ka", 3), ("Tf", 3)]
        rnd.shuffle(keys)
        for k, c in keys[:rnd.randrange(2, len(keys))]:
            mtl.append(f"{k}{ws()}{rnd.randrange(0, 9)}" if c == 0 else k + ws() + ws().join(num() for _ in range(c)))
        if rnd.random() < 0.3:
            mtl.append(f"Tr {num()}")
        mtl.append(rnd.choice(["", "# comment"]))
    L = ["mtllib" + gap() + rnd.choice(["m.mtl", "m.mtl", "m.mtl", "missing.mtl m.mtl", "m.mtl other.mtl", "missing.mtl", "empty.mtl m.mtl"]) + eol()]
    L += ["v" + ws() + ws().join(num() for _ in range(3)) + (ws() + num() if rnd.random() < 0.1 else "") + eol() for _ in range(nv)]
    L += ["vt" + ws() + ws().join(num() for _ in range(rnd.choice([2, 2, 3]))) + eol() for _ in range(nvt)]
    L += ["vn" + ws() + ws().join(num() for _ in range(3)) + eol() for _ in range(nvn)]
    L += [x + eol() for x in ("g group1", "o obj", "s 1") if rnd.random() < 0.4]
    for f in range(rnd.randrange(2, 8)):
        if rnd.random() < 0.6:
            L.append("usemtl" + gap() + rnd.choice(names + ["nope", ""]) + eol())
        form, toks = rnd.randrange(4), []
        for i in rnd.sample(range(1, nv + 1), rnd.choice([3, 3, 3, 4, 5])):
            vi = i if rnd.random() < 0.8 else i - nv - 1
            if form == 1 and nvt:
                toks.append(f"{vi}/{rnd.randrange(1, nvt + 1)}")
            elif form == 2 and nvn:
                toks.append(f"{vi}//{rnd.randrange(1, nvn + 1)}")
            elif form == 3 and nvt and nvn:
                toks.append(f"{vi}/{rnd.randrange(1, nvt + 1)}/{rnd.randrange(1, nvn + 1)}")
            else:
                toks.append(f"{vi}")
        L.append("f" + ws() + ws().join(toks) + eol())
    with open(os.path.join(d, "m.mtl"), "w", newline="") as fh:
        fh.write(rnd.choice(["\n", "\r\n", "\r"]).join(mtl))
    with open(os.path.join(d, "t.obj"), "w", newline="") as fh:
        fh.write("".join(L))
    return os.path.join(d, "t.obj")


def test_obj_mtl_parsing_fuzz_matches_reference(refmod, tmp_path):
    """120 random OBJ/MTL pairs through the reference's loader (tinyobjloader + Scene.cpp + init_materials) and ours:
    triangle records and material records (texture indices included) must be byte-identical, and a file one side
    refuses the other must refuse."""
    import random
    PIL = pytest.importorskip("PIL.Image")
    import adypt_b200 as A
    d = str(tmp_path)
    for nm in ("tex.png", "my tex.png"):
        PIL.fromarray(np.random.default_rng(5).integers(0, 256, size=(4, 4, 3), dtype=np.uint8)).save(os.path.join(d, nm), "PNG")
    open(os.path.join(d, "empty.mtl"), "w").write("")
    open(os.path.join(d, "other.mtl"), "w").write("newmtl zzz\nKd 9 9 9\n")
    rnd = random.Random(20261017)
    textured = 0
    for trial in range(120):
        p = _fuzz_obj(rnd, d)
        try:
            b = refmod.build(p, cache=False)
        except Exception:
            b = None
        try:
            hs = host.HostScene.from_obj(p)
            hs.load_textures()
        except A.AdyptError:
            hs = None
        assert (b is None) == (hs is None), trial
        if b is None:
            continue
        assert hs.tris.shape == b.tris.shape and np.array_equal(hs.tris.view(np.uint8), b.tris.view(np.uint8)), trial
        assert np.array_equal(hs.mats, b.mats), trial
        textured += int((b.mats[:, 0:4].copy().view(np.int32) >= 0).any()) if b.mats.size else 0
    assert textured >= 5
