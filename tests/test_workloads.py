"""Synthetic workloads are deterministic and have the sizes BASELINE.json names."""
import numpy as np

from adypt_b200 import workloads as W
from conftest import fnv1a


def test_sphere_lattice_is_c1():
    m = W.sphere_lattice(5)
    assert m.n_tris == 65536
    assert fnv1a(m.verts) == fnv1a(W.sphere_lattice(5).verts)


def test_city_sizes_and_determinism():
    a, b = W.city(40, 1), W.city(40, 1)
    assert np.array_equal(a.verts, b.verts) and np.array_equal(a.faces, b.faces)
    assert not np.array_equal(a.verts, W.city(40, 2).verts)
    assert a.n_tris % 12 == 2  # boxes + ground quad
    # C2 is ~1.0M triangles (183 cells), C4 ~10M (577 cells): expected box count = cells^2 * 2.5
    assert abs(183 * 183 * 2.5 * 12 - 1.0e6) < 0.02e6
    assert abs(577 * 577 * 2.5 * 12 - 10.0e6) < 0.02e6


def test_obj_round_trip_is_exact(tmp_path, refmod):
    m = W.city(6, 9, mixed_materials=True)
    b = refmod.build(m.write_obj(str(tmp_path)), cache=False)
    assert np.array_equal(b.positions(), m.positions())  # %.9g floats survive tinyobj's parser bit for bit
    matid = b.tris[:, 96:100].copy().view(np.int32).ravel()
    assert np.array_equal(matid, m.face_mat)
    illum = b.mats[:, 48:52].copy().view(np.int32).ravel()
    assert illum.tolist() == [x.illum for x in m.materials]


def test_bounce_rays_are_deterministic_and_unit():
    m = W.tiny_scene("strip")
    prim = np.zeros((4, 8), dtype=np.float32)
    prim[:, :3] = [[0.5, 0.5, 3], [2.5, 0.5, 3], [4.5, 0.25, 3], [100, 100, 3]]
    prim[:, 6] = -1
    tri = np.array([0, 2, 4, -1], dtype=np.int32)
    uv = np.array([[0.2, 0.3]] * 4, dtype=np.float32)
    a = W.bounce_rays(m.positions(), prim, tri, uv, per_hit=8)
    b = W.bounce_rays(m.positions(), prim, tri, uv, per_hit=8)
    assert a.shape == (24, 8) and np.array_equal(a, b)
    assert np.allclose(np.linalg.norm(a[:, 4:7], axis=1), 1.0, atol=1e-6)
    assert (a[:, 6] > 0).all()  # hemisphere about the normal facing the incoming ray (+z)
    assert not (a[:, 4:7] == 0).any()
