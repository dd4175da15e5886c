"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_build/liboracle.so, the CPU restatement of the hot
path (oracle.cpp). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; nothing under adypt_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


class PTConfig(C.Structure):
    _fields_ = [("max_bounce", C.c_int32), ("subpixel", C.c_int32), ("tmp_lifetime", C.c_int32),
                ("ray_tmin", C.c_float), ("clamp", C.c_float), ("sun", C.c_float * 3)]


def build():
    subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        l = C.CDLL(_LIB_PATH)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        l.oracle_trace_closest.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, i32]
        l.oracle_trace_any.argtypes = [vp, vp, vp, u64, vp, vp, i32]
        l.oracle_brute_closest.argtypes = [vp, vp, C.c_uint32, vp, u64, vp, vp, vp, i32]
        l.oracle_mat4_inverse.argtypes = [vp, vp]
        l.oracle_build_woop.argtypes = [vp, vp, C.c_uint32, vp]
        l.oracle_camera_matrices.argtypes = [C.c_float, C.c_float, C.c_float, i32, i32, vp, vp, vp, vp]
        l.oracle_primary_rays.argtypes = [vp, vp, vp, i32, i32, C.c_float, C.c_float, vp]
        l.oracle_sobol_directions.argtypes = [C.c_uint32, vp]
        l.oracle_sobol_sequence.argtypes = [C.c_uint32, C.c_uint32, vp]
        l.oracle_sobol_at.argtypes = [C.c_uint32, C.c_uint32, vp]
        l.oracle_pt_render.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, C.POINTER(PTConfig), vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32]
        l.oracle_primary_view.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, vp]
        l.oracle_det_sincos.argtypes = [vp, u64, vp, vp]
        l.oracle_det_pow.argtypes = [vp, vp, u64, vp]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def hardware_threads() -> int:
    return int(lib().oracle_hardware_threads())


def trace_closest(nodes, tri_indices, woop, rays, nthreads=0, want_t=True):
    """-> dict(tri (n,) i32 scene ids, t (n,) f32, uv (n,2) f32, counters dict)."""
    nodes, tri_indices, woop, rays = _c(nodes, np.uint8), _c(tri_indices, np.int32), _c(woop, np.float32), _c(rays, np.float32)
    n = rays.size // 8
    tri = np.empty(n, dtype=np.int32)
    t = np.empty(n, dtype=np.float32) if want_t else None
    uv = np.empty((n, 2), dtype=np.float32)
    cnt = np.zeros(4, dtype=np.uint64)
    lib().oracle_trace_closest(_p(nodes), _p(tri_indices), _p(woop), _p(rays), n, _p(tri), _p(t), _p(uv), _p(cnt), nthreads)
    return dict(tri=tri, t=t, uv=uv, counters=dict(nodes=int(cnt[0]), tris=int(cnt[1]), max_stack=int(cnt[2]), hits=int(cnt[3])))


def trace_any(nodes, woop, rays, nthreads=0):
    nodes, woop, rays = _c(nodes, np.uint8), _c(woop, np.float32), _c(rays, np.float32)
    n = rays.size // 8
    occ = np.empty(n, dtype=np.uint8)
    cnt = np.zeros(4, dtype=np.uint64)
    lib().oracle_trace_any(_p(nodes), _p(woop), _p(rays), n, _p(occ), _p(cnt), nthreads)
    return dict(occluded=occ, counters=dict(nodes=int(cnt[0]), tris=int(cnt[1]), max_stack=int(cnt[2]), hits=int(cnt[3])))


def brute_closest(tri_indices, woop, rays, nthreads=0):
    tri_indices, woop, rays = _c(tri_indices, np.int32), _c(woop, np.float32), _c(rays, np.float32)
    n = rays.size // 8
    tri = np.empty(n, dtype=np.int32)
    t = np.empty(n, dtype=np.float32)
    uv = np.empty((n, 2), dtype=np.float32)
    lib().oracle_brute_closest(_p(tri_indices), _p(woop), tri_indices.size, _p(rays), n, _p(tri), _p(t), _p(uv), nthreads)
    return dict(tri=tri, t=t, uv=uv)


def mat4_inverse(m):
    a = _c(m, np.float32).reshape(16)
    o = np.zeros(16, dtype=np.float32)
    lib().oracle_mat4_inverse(_p(a), _p(o))
    return o


def build_woop(tris100, tri_indices):
    tris100, tri_indices = _c(tris100, np.uint8), _c(tri_indices, np.int32)
    out = np.zeros((tri_indices.size, 12), dtype=np.float32)
    lib().oracle_build_woop(_p(tris100), _p(tri_indices), tri_indices.size, _p(out))
    return out


def camera_matrices(fov, yaw, pitch, width, height):
    ms = [np.zeros(16, dtype=np.float32) for _ in range(4)]
    lib().oracle_camera_matrices(fov, yaw, pitch, width, height, *[_p(m) for m in ms])
    return dict(proj=ms[0], view=ms[1], inv_proj=ms[2], inv_view=ms[3])


def primary_rays(origin, tmin, inv_proj, inv_view, width, height, bias=(0.0, 0.0)):
    ot = np.array([origin[0], origin[1], origin[2], tmin], dtype=np.float32)
    ip, iv = _c(inv_proj, np.float32), _c(inv_view, np.float32)
    rays = np.zeros((width * height, 8), dtype=np.float32)
    lib().oracle_primary_rays(_p(ot), _p(ip), _p(iv), width, height, bias[0], bias[1], _p(rays))
    return rays


def sobol_directions(dim_index):
    o = np.zeros(32, dtype=np.uint32)
    if lib().oracle_sobol_directions(dim_index, _p(o)) != 0:
        raise ValueError("dimension out of range")
    return o


def sobol_sequence(dim, n_calls):
    o = np.zeros((n_calls, dim), dtype=np.float32)
    if lib().oracle_sobol_sequence(dim, n_calls, _p(o)) != 0:
        raise ValueError("dimension out of range")
    return o


def sobol_at(dim, index):
    o = np.zeros(dim, dtype=np.float32)
    if lib().oracle_sobol_at(dim, index, _p(o)) != 0:
        raise ValueError("dimension out of range")
    return o


def _textures(textures):
    """list of (h,w,3) uint8 arrays -> (n, pointer array, wh array, keepalive)"""
    if not textures:
        return 0, None, None, None
    arrs = [_c(t, np.uint8) for t in textures]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    wh = np.array([[a.shape[1], a.shape[0]] for a in arrs], dtype=np.int32)
    return len(arrs), ptrs, wh, arrs


def pt_render(bvh, origin, inv_proj, inv_view, width, height, cfg: dict, bias_rg8, first_spp, n_spp,
              out_rgba=None, primary_tmp=None, nthreads=0, sum_mode=False, textures=None, sun_visibility=None, russian_roulette=None):
    """bvh: object with nodes/tri_indices/woop/tris/mats arrays. Returns (out_rgba, primary_tmp, counters)."""
    c = PTConfig(cfg["max_bounce"], cfg["subpixel"], cfg["tmp_lifetime"], cfg["ray_tmin"], cfg["clamp"],
                 (C.c_float * 3)(*cfg["sun"]))
    npix = width * height
    if out_rgba is None:
        out_rgba = np.zeros((npix, 4), dtype=np.float32)
    if primary_tmp is None:
        primary_tmp = np.zeros((npix, 4), dtype=np.float32)
    o = np.array(origin, dtype=np.float32)
    ip, iv = _c(inv_proj, np.float32), _c(inv_view, np.float32)
    bias = _c(bias_rg8, np.uint8)
    cnt = np.zeros(5, dtype=np.uint64)
    nt, tp, twh, _keep = _textures(textures)
    sv = None if sun_visibility is None else np.ascontiguousarray(sun_visibility, dtype=np.float32)
    nodes, ti, woop = _c(bvh.nodes, np.uint8), _c(bvh.tri_indices, np.int32), _c(bvh.woop, np.float32)
    tris, mats = _c(bvh.tris, np.uint8), _c(bvh.mats, np.uint8)
    rc = lib().oracle_pt_render(_p(nodes), _p(ti), _p(woop), _p(tris), _p(mats), _p(o), _p(ip), _p(iv), width, height,
                                C.byref(c), _p(bias), first_spp, n_spp, _p(out_rgba), _p(primary_tmp), _p(cnt), nthreads, int(sum_mode), nt, tp, _p(twh), _p(sv),
                                -1 if russian_roulette is None else int(russian_roulette))
    if rc != 0:
        raise RuntimeError("oracle_pt_render failed")
    return out_rgba, primary_tmp, dict(nodes=int(cnt[0]), tris=int(cnt[1]), max_stack=int(cnt[2]), hits=int(cnt[3]), segments=int(cnt[4]))


def primary_view(bvh, origin, tmin, inv_proj, inv_view, width, height, vtype, nthreads=0, textures=None):
    ot = np.array([origin[0], origin[1], origin[2], tmin], dtype=np.float32)
    ip, iv = _c(inv_proj, np.float32), _c(inv_view, np.float32)
    out = np.zeros((width * height, 4), dtype=np.float32)
    nt, tp, twh, _keep = _textures(textures)
    nodes, ti, woop = _c(bvh.nodes, np.uint8), _c(bvh.tri_indices, np.int32), _c(bvh.woop, np.float32)
    tris, mats = _c(bvh.tris, np.uint8), _c(bvh.mats, np.uint8)
    lib().oracle_primary_view(_p(nodes), _p(ti), _p(woop), _p(tris), _p(mats), _p(ot), _p(ip), _p(iv), width, height, vtype, _p(out), nthreads, nt, tp, _p(twh))
    return out


def det_sincos(x):
    x = _c(x, np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().oracle_det_sincos(_p(x), x.size, _p(s), _p(c))
    return s, c


def det_pow(x, y):
    x, y = _c(x, np.float32), _c(y, np.float32)
    out = np.empty_like(x)
    lib().oracle_det_pow(_p(x), _p(y), x.size, _p(out))
    return out
