"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libadypt_ref.so, the UNMODIFIED reference CPU
pipeline (OBJ -> Triangle[] -> SBVH -> CWBVH, Sobol, .config) compiled in place from /root/reference by
oracle/Makefile. Only tests/, __graft_entry__.smoke() and bench.py's input preparation / cpu_baseline leg
use this; nothing under adypt_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libadypt_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{_LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        l = C.CDLL(_LIB_PATH)
        l.ref_scene_build.restype = C.c_void_p
        l.ref_scene_build.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_float]
        l.ref_scene_load_bvh.restype = C.c_void_p
        l.ref_scene_load_bvh.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_float, C.c_float]
        l.ref_scene_save_bvh.restype = C.c_int
        l.ref_scene_save_bvh.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_float, C.c_float]
        l.ref_scene_destroy.argtypes = [C.c_void_p]
        for n in ("n_tris", "n_nodes", "n_refs", "n_mats", "n_sbvh_nodes"):
            f = getattr(l, "ref_scene_" + n)
            f.restype = C.c_uint32
            f.argtypes = [C.c_void_p]
        for n in ("tris", "nodes", "tri_indices", "woop", "mats", "sbvh_nodes"):
            f = getattr(l, "ref_scene_" + n)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p]
        l.ref_scene_aabb.argtypes = [C.c_void_p, C.c_void_p]
        l.ref_mat4_inverse.argtypes = [C.c_void_p, C.c_void_p]
        l.ref_sobol_sequence.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        l.ref_camera_matrices.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 4
        l.ref_config_load.restype = C.c_int
        l.ref_config_load.argtypes = [C.c_char_p, C.c_void_p]
        l.ref_config_roundtrip_json.restype = C.c_int
        l.ref_config_roundtrip_json.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32]
        l.ref_save_exr.restype = C.c_int
        l.ref_save_exr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        l.ref_load_image_rgb8.restype = C.c_int
        l.ref_load_image_rgb8.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        _lib = l
    return _lib


def _copy(ptr, nbytes, dtype):
    if nbytes == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class RefBVH:
    """Arrays produced by the reference pipeline, copied into numpy (the GPU ABI of SURVEY.md §8a)."""

    def __init__(self, h, l):
        nt, nn, nr, nm = l.ref_scene_n_tris(h), l.ref_scene_n_nodes(h), l.ref_scene_n_refs(h), l.ref_scene_n_mats(h)
        self.tris = _copy(l.ref_scene_tris(h), nt * 100, np.uint8).reshape(nt, 100)
        self.nodes = _copy(l.ref_scene_nodes(h), nn * 80, np.uint8).reshape(nn, 80)
        self.tri_indices = _copy(l.ref_scene_tri_indices(h), nr * 4, np.int32)
        self.woop = _copy(l.ref_scene_woop(h), nr * 48, np.float32).reshape(nr, 12)
        self.mats = _copy(l.ref_scene_mats(h), nm * 64, np.uint8).reshape(nm, 64)
        ns = l.ref_scene_n_sbvh_nodes(h)
        self.sbvh_nodes = _copy(l.ref_scene_sbvh_nodes(h), ns * 32, np.uint8).reshape(ns, 32) if ns else None
        a = np.zeros(6, dtype=np.float32)
        l.ref_scene_aabb(h, a.ctypes.data)
        self.aabb = a

    @property
    def n_tris(self):
        return self.tris.shape[0]

    @property
    def n_nodes(self):
        return self.nodes.shape[0]

    @property
    def n_refs(self):
        return self.tri_indices.shape[0]

    def positions(self):
        """(n_tris,3,3) f32 corner positions out of the 100-byte Triangle records."""
        return self.tris[:, :36].copy().view(np.float32).reshape(-1, 3, 3)

    def save(self, path):
        np.savez(path, tris=self.tris, nodes=self.nodes, tri_indices=self.tri_indices, woop=self.woop,
                 mats=self.mats, aabb=self.aabb)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        o = cls.__new__(cls)
        o.tris, o.nodes, o.tri_indices, o.woop, o.mats, o.aabb = (z[k] for k in ("tris", "nodes", "tri_indices", "woop", "mats", "aabb"))
        o.sbvh_nodes = None
        return o


def build(obj_path: str, max_spatial_depth: int = 48, triangle_sah: float = 0.3, node_sah: float = 1.0,
          cache: bool = True) -> RefBVH:
    """Reference OBJ -> CWBVH build (Instance.cpp:12-24). Cached as <obj>.refbvh.npz."""
    tag = f".refbvh_{max_spatial_depth}_{triangle_sah:g}_{node_sah:g}.npz"
    cpath = obj_path + tag
    if cache and os.path.exists(cpath):
        return RefBVH.load(cpath)
    l = lib()
    h = l.ref_scene_build(obj_path.encode(), max_spatial_depth, triangle_sah, node_sah)
    if not h:
        raise RuntimeError(f"reference failed to load {obj_path}")
    try:
        r = RefBVH(h, l)
    finally:
        l.ref_scene_destroy(h)
    if cache:
        r.save(cpath)
    return r


def build_keep_handle(obj_path: str, max_spatial_depth=48, triangle_sah=0.3, node_sah=1.0):
    l = lib()
    h = l.ref_scene_build(obj_path.encode(), max_spatial_depth, triangle_sah, node_sah)
    if not h:
        raise RuntimeError(f"reference failed to load {obj_path}")
    return h


def mat4_inverse(m: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    o = np.zeros(16, dtype=np.float32)
    lib().ref_mat4_inverse(a.ctypes.data, o.ctypes.data)
    return o


def sobol_sequence(dim: int, n_calls: int) -> np.ndarray:
    o = np.zeros((n_calls, dim), dtype=np.float32)
    lib().ref_sobol_sequence(dim, n_calls, o.ctypes.data)
    return o


def camera_matrices(fov, yaw, pitch, width, height):
    ms = [np.zeros(16, dtype=np.float32) for _ in range(4)]
    lib().ref_camera_matrices(fov, yaw, pitch, width, height, *[m.ctypes.data for m in ms])
    return dict(proj=ms[0], view=ms[1], inv_proj=ms[2], inv_view=ms[3])


class RefConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("invocation_size", C.c_int32), ("stack_size", C.c_int32), ("max_bounce", C.c_int32),
        ("subpixel", C.c_int32), ("tmp_lifetime", C.c_int32),
        ("ray_tmin", C.c_float), ("clamp", C.c_float), ("sun", C.c_float * 3),
        ("max_spatial_depth", C.c_int32), ("triangle_sah", C.c_float), ("node_sah", C.c_float),
        ("speed", C.c_float), ("mouse_sensitive", C.c_float), ("fov", C.c_float), ("yaw", C.c_float),
        ("pitch", C.c_float), ("position", C.c_float * 3),
        ("obj_filename", C.c_char * 512), ("bvh_filename", C.c_char * 512),
    ]


def config_load(path: str):
    c = RefConfig()
    rc = lib().ref_config_load(path.encode(), C.byref(c))
    return c if rc == 0 else None


def config_roundtrip_json(path: str):
    buf = C.create_string_buffer(1 << 16)
    n = lib().ref_config_roundtrip_json(path.encode(), buf, len(buf))
    return buf.value.decode() if n >= 0 else None


def dtoa(values):
    """rapidjson's own double formatting (what its Writer emits) for an array of doubles -> list of str."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    buf = C.create_string_buffer(32 * max(1, v.size))
    l = lib()
    l.ref_dtoa.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    l.ref_dtoa(v.ctypes.data, v.size, buf)
    raw = buf.raw
    return [raw[32 * i: 32 * i + 32].split(b"\0", 1)[0].decode() for i in range(v.size)]


def save_exr(rgb, path: str, fp16: bool = False) -> int:
    """SaveEXR of the reference's vendored tinyexr, called like OglPathTracer::SaveResult does; rgb: (h, w, 3) float32."""
    a = np.ascontiguousarray(rgb, dtype=np.float32)
    return lib().ref_save_exr(a.ctypes.data, a.shape[1], a.shape[0], int(fp16), path.encode())


def load_image_rgb8(path: str):
    """stbi_load(path, ..., 3) of the reference's vendored stb_image -> (h,w,3) uint8 array, or None on failure."""
    w, h = C.c_int(0), C.c_int(0)
    buf = np.zeros(1 << 26, dtype=np.uint8)
    rc = lib().ref_load_image_rgb8(path.encode(), C.byref(w), C.byref(h), buf.ctypes.data, buf.size)
    if rc != 0:
        return None
    return buf[: w.value * h.value * 3].reshape(h.value, w.value, 3).copy()
