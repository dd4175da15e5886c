"""CUDA kernels against answers computed by the REFERENCE'S OWN SHADERS: tests/golden/glsl_city12.npz was produced by
shaders/traversal.glsl / primaryray.glsl / pathtracer.glsl running on the CPU (oracle/glsl_transpile.py) with the
reference's Sobol generator -- see tests/golden/make_glsl_golden.py. No oracle in this chain; everything bit for bit."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN, "glsl_city12.npz")), load_golden("city12")


def test_traversal_equals_reference_shader(A, fx):
    z, g = fx
    sc = A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
    rays = g.extra["rays"]
    got = sc.trace_closest(rays)
    hit = z["glsl_tri"] >= 0
    assert hit.any() and (~hit).any()
    assert np.array_equal(got["tri"], z["glsl_tri"])
    assert np.array_equal(bits(got["uv"])[hit], bits(z["glsl_uv"])[hit])
    assert np.array_equal(sc.trace_any(rays), z["glsl_any"])


def test_viewer_images_equal_reference_shader(A, fx):
    z, g = fx
    w, h = (int(v) for v in z["size"])
    sc = A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
    tr = A.Tracer(sc, A.PTConfig.make(), w, h, bias_seed=1)
    cam = g.extra["cam"]
    tr.look(cam[:3], float(cam[3]), float(cam[4]), float(cam[5]))
    for vt in (0, 1, 2, 4, 5):
        tr.primary(vt)
        assert np.array_equal(bits(tr.read(4).reshape(-1, 4)), bits(z[f"view{vt}"])), vt


@pytest.mark.parametrize("tag", ["a", "b"])
def test_path_traced_image_equals_reference_shader(A, fx, tag):
    z, g = fx
    w, h = (int(v) for v in z["size"])
    c = z["cfg_" + tag]
    sc = A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
    pc = A.PTConfig.make(max_bounce=int(c[0]), subpixel=int(c[1]), tmp_lifetime=int(c[2]), ray_tmin=float(np.float32(c[3])),
                         clamp=float(c[4]), sun=tuple(float(v) for v in c[5:8]))
    tr = A.Tracer(sc, pc, w, h, bias_seed=1)
    tr.set_bias(z["bias"])  # uSobolBiasImg
    cam = g.extra["cam"]
    tr.look(cam[:3], float(cam[3]), float(cam[4]), float(cam[5]))
    tr.sample(int(c[8]))
    img = tr.read(4).reshape(-1, 4)
    assert np.array_equal(bits(img), bits(z["pt_" + tag]))
    assert img[:, :3].mean() > 0.05
