"""C-ABI library checks that need no GPU: it loads, exports every symbol the header declares, fails loudly
(no CPU fallback) without a device, and its host-side arithmetic / EXR writer are right."""
import ctypes as C
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, gpu_available, load_golden

import adypt_b200 as A
from exr_reader import read_exr


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_library_is_built_in_tree():
    assert os.path.exists(A.LIB_PATH), "run python -c 'import __graft_entry__ as g; g.build()'"
    assert A.LIB_PATH.startswith(ROOT)


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "adypt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(adypt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    lib = A.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/adypt_b200.h but not exported"
    assert sorted(A.EXPORTS) == declared
    assert lib.adypt_version() == 100


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "adypt_b200.h"\nint main(void){ adypt_pt_config c; return sizeof(c) == 40 && sizeof(adypt_tracer_profile) == %d ? 0 : 1; }\n'
                   % C.sizeof(A.TracerProfile))  # the ctypes mirror of the profile block has the C layout
    exe = tmp_path / "t"
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


@pytest.mark.skipif(gpu_available(), reason="checks behaviour on a machine WITHOUT a GPU")
def test_no_device_fails_loudly():
    g = load_golden("tiny_strip")
    with pytest.raises(A.AdyptError) as e:
        A.Scene(g.nodes, g.tri_indices, g.woop, g.tris, g.mats)
    assert e.value.code == -2  # ADYPT_ENODEV
    assert "no CPU fallback" in str(e.value) or "cudaGetDeviceCount" in str(e.value)


def test_argument_validation():
    lib = A.load_library()
    assert lib.adypt_scene_create(None, None) == -1
    assert b"NULL" in lib.adypt_last_error()
    assert lib.adypt_trace_closest(None, None, 0, None, None, None, 0, None) == -1
    assert lib.adypt_tracer_sample(None, 1) == -1
    assert lib.adypt_scene_destroy(None) == 0 and lib.adypt_tracer_destroy(None) == 0


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    paths = glob.glob(os.path.join(ROOT, "adypt_b200", "**", "*"), recursive=True) + glob.glob(os.path.join(ROOT, "tools", "*.py"))
    for path in paths:
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
            txt = open(path, errors="replace").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), path
            assert "liboracle" not in txt and "libadypt_ref" not in txt, path
            assert not re.search(r'#include\s+"[^"]*oracle', txt), path


def test_camera_matrices_match_reference_golden():
    """adypt_camera_matrices == Camera::GetProjection/GetView of the reference (golden from oracle/_ref)."""
    z = np.load(os.path.join(GOLDEN, "ref_camera.npz"))
    for i, (fov, yaw, pitch, w, h) in enumerate(z["params"]):
        p, v = A.camera_matrices(float(fov), float(yaw), float(pitch), int(w), int(h))
        assert np.array_equal(bits(p), bits(z["proj"][i])), i
        assert np.array_equal(bits(v), bits(z["view"][i])), i


def test_product_sobol_equals_reference_golden():
    """The product's own decode of the Joe-Kuo table + closed-form generator (csrc/hostmath.cpp) against the reference's
    Sobol::Next vectors: dims 1-64 and 9901-10005 value by value over 300 calls (no GPU, no oracle)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_sobol.npz"))
    for idx in list(range(0, 300, 7)) + [299]:
        v = A.sobol_vector(10005, idx)
        assert np.array_equal(v[:64].view(np.uint32), z["seq"][idx].view(np.uint32)), idx
        assert np.array_equal(v[9900:].view(np.uint32), z["tail"][idx].view(np.uint32)), idx
    with pytest.raises(A.AdyptError):
        A.sobol_vector(10006, 0)


@pytest.mark.parametrize("fp16", [False, True])
def test_exr_round_trip(tmp_path, fp16):
    rng = np.random.default_rng(3)
    h, w = 37, 53  # not a multiple of the 16-line ZIP chunk
    img = rng.random((h, w, 3), dtype=np.float32) * 4.0
    img[0, 0] = [0.0, 1.0, 65504.0]
    img[1, 1] = [1e-8, 6e-5, 70000.0]  # half: flush / subnormal / overflow to inf
    path = str(tmp_path / "o.exr")
    A.write_exr(path, img, fp16=fp16)
    r = read_exr(path)
    assert (r["width"], r["height"]) == (w, h)
    assert [c for c, _ in r["channels"]] == ["B", "G", "R"]
    assert r["compression"] == 3  # ZIP, like tinyexr's SaveEXR default
    got = np.stack([r["data"]["R"], r["data"]["G"], r["data"]["B"]], axis=2)
    if fp16:
        # tinyexr's conversion rounds halves up: never further than one half-ulp step from numpy's ties-to-even
        near = img.astype(np.float16)
        assert np.array_equal(got[1:], near[1:].astype(np.float32)) or np.all(np.abs(got.astype(np.float16).view(np.int16).astype(int) - near.view(np.int16).astype(int)) <= 1)
    else:
        assert np.array_equal(bits(got), bits(img))


@pytest.mark.parametrize("fp16", [False, True])
def test_exr_equals_the_reference_writer(refmod, tmp_path, fp16):
    """The same image through the reference's own tinyexr SaveEXR (called as OglPathTracer::SaveResult calls it) and
    through adypt_write_exr: same header attributes, same channel layout, same pixel bits -- including tinyexr's float
    -> half rules (halves round up, float subnormals flush, NaN -> 0x7e00, overflow -> inf). Only the deflate streams
    differ (miniz there, zlib here)."""
    rng = np.random.default_rng(1)
    h, w = 37, 53
    img = (rng.random((h, w, 3), dtype=np.float32) * 4).astype(np.float32)
    img[0, 0] = [0.0, 1.0, 65504.0]
    img[1, 1] = [1e-8, 6e-5, 70000.0]
    img[2, 2] = [np.inf, -np.inf, np.nan]
    img[3, 3] = [-0.0, -1e-8, 6.1e-5]
    img[4, 4] = [65519.9, 65520.0, -65520.0]
    img[5, 5] = [5.96e-8, 2.98e-8, 2.99e-8]
    img[6] = (rng.random((w, 3)) * 1e-6).astype(np.float32)                                      # subnormal halves
    img[7, :, 0] = np.float32(2.0) ** -14 * (1 - rng.random(w).astype(np.float32) * 2 ** -10)      # around the smallest normal half
    img[8, :, 1] = (np.arange(w, dtype=np.float32) + np.float32(0.5)) * np.float32(2.0 ** -11) + 1  # exact ties
    img[9, :, 2] = np.array([1e-40, -1e-39, 1e-45], dtype=np.float32).repeat(18)[:w]                # float subnormals
    a, b = str(tmp_path / "ref.exr"), str(tmp_path / "own.exr")
    assert refmod.save_exr(img, a, fp16) == 0
    A.write_exr(b, img, fp16=fp16)
    r, o = read_exr(a), read_exr(b)
    assert r["channels"] == o["channels"] and r["attrs"] == o["attrs"]
    for c in "RGB":
        x, y = r["data"][c], o["data"][c]
        assert np.array_equal(np.isnan(x), np.isnan(y)), c
        assert np.array_equal(bits(x)[~np.isnan(x)], bits(y)[~np.isnan(y)]), c


@pytest.mark.skipif(not os.path.isdir("/root/reference/results"), reason="reference tree not present")
def test_reader_parses_reference_showcase_exr():
    """Pins the test reader (and with it the writer's layout) on a file written by the reference's tinyexr."""
    f = sorted(glob.glob("/root/reference/results/*.exr"))[0]
    r = read_exr(f)
    assert (r["width"], r["height"]) == (1280, 720)
    assert [c for c, _ in r["channels"]] == ["B", "G", "R"] and all(t == 1 for _, t in r["channels"])
    assert r["compression"] == 3
    assert np.isfinite(r["data"]["R"]).all() and r["data"]["R"].mean() > 0


def test_no_exception_crosses_the_c_boundary():
    """Every extern "C" body runs inside guarded(): a std::bad_alloc inside the library comes back as ADYPT_ENOMEM
    with a message instead of terminating the host process. Provoked in a child process with a 4 GiB address-space
    limit and a 200 M-triangle request (the 20 GB vector resize fails before any input is read)."""
    code = r"""
import ctypes as C, resource, sys
sys.path.insert(0, %r)
import numpy as np
import adypt_b200 as A
from adypt_b200 import host
lib = host._lib()
pos = np.zeros((4, 9), dtype=np.float32); mid = np.zeros(4, dtype=np.int32); mats = np.zeros((1, 64), dtype=np.uint8)
resource.setrlimit(resource.RLIMIT_AS, (4 << 30, 4 << 30))
h = C.c_void_p()
rc = lib.adypt_host_scene_from_triangles(pos.ctypes.data, mid.ctypes.data, 200_000_000, mats.ctypes.data, 1, C.byref(h))
print("RC", rc, A.load_library().adypt_last_error().decode())
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "RC -4 out of host memory" in r.stdout, r.stdout + r.stderr[-500:]
