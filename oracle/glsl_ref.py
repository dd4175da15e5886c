"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/libadypt_glsl.so -- the reference's compute shaders
(shaders/traversal.glsl, primaryray.glsl, pathtracer.glsl) compiled for the CPU from the shader text where it lies
(oracle/glsl_transpile.py + glsl_shim.inc + glsl_driver.inc, built by oracle/Makefile). It pins the oracle's
traversal and shading to the reference's own code; nothing under adypt_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libadypt_glsl.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{_LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        l = C.CDLL(_LIB_PATH)
        vp, i32, u64, f32 = C.c_void_p, C.c_int32, C.c_uint64, C.c_float
        l.glsl_trace_closest.argtypes = [vp, vp, vp, vp, u64, vp, vp, i32]
        l.glsl_trace_any.argtypes = [vp, vp, vp, u64, vp, i32]
        l.glsl_pt_render.argtypes = [vp] * 8 + [i32, i32, i32, i32, i32, f32, vp, vp, vp, i32, i32, vp, vp, i32, vp, vp, i32]
        l.glsl_primary_view.argtypes = [vp] * 8 + [i32, i32, i32, vp, i32, vp, vp, i32]
        _lib = l
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return None if a is None else a.ctypes.data


def _textures(textures):
    if not textures:
        return 0, None, None, []
    arrs = [_c(t, np.uint8) for t in textures]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    wh = np.array([[a.shape[1], a.shape[0]] for a in arrs], dtype=np.int32)
    return len(arrs), ptrs, wh, arrs


def trace_closest(nodes, tri_indices, woop, rays, nthreads=0):
    """BVHIntersection(origin_tmin, dir, out idx, inout uv) of traversal.glsl for every ray -> (tri, uv)."""
    nodes, ti, woop, rays = _c(nodes, np.uint8), _c(tri_indices, np.int32), _c(woop, np.float32), _c(rays, np.float32)
    n = rays.shape[0]
    tri, uv = np.empty(n, dtype=np.int32), np.empty((n, 2), dtype=np.float32)
    lib().glsl_trace_closest(_p(nodes), _p(ti), _p(woop), _p(rays), n, _p(tri), _p(uv), nthreads)
    return tri, uv


def trace_any(nodes, woop, rays, nthreads=0):
    nodes, woop, rays = _c(nodes, np.uint8), _c(woop, np.float32), _c(rays, np.float32)
    occ = np.empty(rays.shape[0], dtype=np.uint8)
    lib().glsl_trace_any(_p(nodes), _p(woop), _p(rays), rays.shape[0], _p(occ), nthreads)
    return occ


def pt_render(bvh, origin, inv_proj, inv_view, width, height, cfg: dict, bias_rg8, sobol, first_spp, n_spp,
              out_rgba=None, primary_tmp=None, nthreads=0, textures=None):
    """n_spp dispatches of pathtracer.glsl. `sobol`: float32 [frames][2*max_bounce], the reference's Sobol::Next
    vectors from frame 0 on. Returns (out_rgba, primary_tmp)."""
    npix = width * height
    if out_rgba is None:
        out_rgba = np.zeros((npix, 4), dtype=np.float32)
    if primary_tmp is None:
        primary_tmp = np.zeros((npix, 4), dtype=np.float32)
    ot = np.array([origin[0], origin[1], origin[2], cfg["ray_tmin"]], dtype=np.float32)
    ip, iv = _c(inv_proj, np.float32), _c(inv_view, np.float32)
    bias, sob, sun = _c(bias_rg8, np.uint8), _c(sobol, np.float32), _c(cfg["sun"], np.float32)
    assert sob.shape[0] >= first_spp + n_spp and sob.shape[1] == 2 * cfg["max_bounce"]
    nt, tp, twh, _keep = _textures(textures)
    nodes, ti, woop = _c(bvh.nodes, np.uint8), _c(bvh.tri_indices, np.int32), _c(bvh.woop, np.float32)
    tris, mats = _c(bvh.tris, np.uint8), _c(bvh.mats, np.uint8)
    lib().glsl_pt_render(_p(nodes), _p(ti), _p(woop), _p(tris), _p(mats), _p(ot), _p(ip), _p(iv), width, height,
                         cfg["max_bounce"], cfg["subpixel"], cfg["tmp_lifetime"], cfg["clamp"], _p(sun), _p(bias), _p(sob),
                         first_spp, n_spp, _p(out_rgba), _p(primary_tmp), nt, tp, _p(twh), nthreads)
    return out_rgba, primary_tmp


def primary_view(bvh, origin, tmin, inv_proj, inv_view, width, height, vtype, nthreads=0, textures=None):
    ot = np.array([origin[0], origin[1], origin[2], tmin], dtype=np.float32)
    ip, iv = _c(inv_proj, np.float32), _c(inv_view, np.float32)
    out = np.zeros((width * height, 4), dtype=np.float32)
    nt, tp, twh, _keep = _textures(textures)
    nodes, ti, woop = _c(bvh.nodes, np.uint8), _c(bvh.tri_indices, np.int32), _c(bvh.woop, np.float32)
    tris, mats = _c(bvh.tris, np.uint8), _c(bvh.mats, np.uint8)
    lib().glsl_primary_view(_p(nodes), _p(ti), _p(woop), _p(tris), _p(mats), _p(ot), _p(ip), _p(iv), width, height, vtype, _p(out),
                            nt, tp, _p(twh), nthreads)
    return out
