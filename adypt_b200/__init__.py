"""adypt_b200: B200-native (sm_100a) ray-traversal and path-tracing core for AdamYuan/Adypt.

This package is a thin ctypes binding over the C-ABI in ``include/adypt_b200.h`` (built in-tree as
``adypt_b200/lib/libadypt_b200.so``). It mirrors the reference's tracer surface --
``OglScene::Initialize`` -> :class:`Scene`, ``OglPathTracer`` -> :class:`Tracer`
(``src/Tracer/OglScene.hpp:43``, ``src/Tracer/OglPathTracer.hpp:66-82``) -- plus the batch form of
``BVHIntersection`` (``shaders/traversal.glsl:14-255, 257-494``).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is present, every compute
call raises. Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libadypt_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
VIEW_DIFFUSE, VIEW_SPECULAR, VIEW_EMISSIVE, VIEW_RADIANCE, VIEW_NORMAL, VIEW_POSITION = range(6)

# every symbol include/adypt_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "adypt_last_error", "adypt_version", "adypt_device_count", "adypt_scene_create", "adypt_scene_destroy", "adypt_scene_set_textures",
    "adypt_host_scene_load_textures", "adypt_host_scene_texture",
    "adypt_scene_read_woop", "adypt_scene_device_bytes", "adypt_trace_closest", "adypt_trace_any", "adypt_trace_stats", "adypt_launch_count",
    "adypt_trace_configure", "adypt_tracer_create", "adypt_tracer_destroy", "adypt_tracer_set_config",
    "adypt_tracer_set_bias", "adypt_tracer_get_bias", "adypt_tracer_set_sun_visibility", "adypt_tracer_set_russian_roulette", "adypt_tracer_set_camera", "adypt_camera_matrices",
    "adypt_tracer_primary", "adypt_tracer_sample", "adypt_tracer_accumulate", "adypt_tracer_sum_buffer",
    "adypt_tracer_clear_sum", "adypt_tracer_resolve_sum", "adypt_tracer_spp", "adypt_tracer_read",
    "adypt_tracer_result_buffer", "adypt_tracer_save_exr", "adypt_tracer_sync", "adypt_tracer_primary_rays",
    "adypt_sobol_vector", "adypt_tracer_stats", "adypt_tracer_set_profiling", "adypt_tracer_get_profile", "adypt_trace_kernel_name", "adypt_write_exr", "adypt_debug_math",
    "adypt_host_scene_load_obj", "adypt_host_scene_from_triangles", "adypt_host_scene_destroy", "adypt_host_scene_build_bvh",
    "adypt_host_scene_load_bvh", "adypt_host_scene_save_bvh", "adypt_host_scene_get", "adypt_host_scene_upload",
    "adypt_config_format_double", "adypt_config_set_default", "adypt_config_load", "adypt_config_to_json", "adypt_config_save",
    "adypt_group_create", "adypt_group_destroy", "adypt_group_set_camera", "adypt_group_set_sun_visibility", "adypt_group_set_russian_roulette", "adypt_group_render",
    "adypt_group_read", "adypt_group_save_exr", "adypt_tracer_stream",
]


class AdyptError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"adypt_b200 error {code}: {msg}")
        self.code = code


class SceneDesc(C.Structure):
    _fields_ = [("device", C.c_int32), ("nodes", C.c_void_p), ("n_nodes", C.c_uint32), ("tri_indices", C.c_void_p),
                ("n_refs", C.c_uint32), ("woop", C.c_void_p), ("triangles", C.c_void_p), ("n_tris", C.c_uint32),
                ("materials", C.c_void_p), ("n_mats", C.c_uint32)]


class TextureDesc(C.Structure):
    _fields_ = [("rgb8", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32)]


class PTConfig(C.Structure):
    """Memory layout of InstanceConfig::PT (src/InstanceConfig.hpp:22-28)."""
    _fields_ = [("invocation_size", C.c_int32), ("stack_size", C.c_int32), ("max_bounce", C.c_int32),
                ("subpixel", C.c_int32), ("tmp_lifetime", C.c_int32), ("ray_tmin", C.c_float), ("clamp", C.c_float),
                ("sun", C.c_float * 3)]

    @classmethod
    def make(cls, max_bounce=5, subpixel=8, tmp_lifetime=16, ray_tmin=1e-4, clamp=4.0, sun=(1.0, 1.0, 1.0),
             invocation_size=8, stack_size=12):
        return cls(invocation_size, stack_size, max_bounce, subpixel, tmp_lifetime, ray_tmin, clamp, (C.c_float * 3)(*sun))


_lib = None


STAGES = ("generate", "trace_primary", "shade_primary", "trace_bounce", "shade_bounce", "accumulate", "connect", "other")


class TracerProfile(C.Structure):  # adypt_tracer_profile
    _fields_ = [("stage_ms", C.c_double * 8), ("stage_launches", C.c_uint64 * 8), ("trace_nodes", C.c_uint64), ("trace_tris", C.c_uint64),
                ("trace_hits", C.c_uint64), ("trace_rays", C.c_uint64), ("trace_max_depth", C.c_uint64),
                ("primary_nodes", C.c_uint64), ("primary_tris", C.c_uint64), ("primary_hits", C.c_uint64), ("primary_rays", C.c_uint64)]


def load_library():
    """dlopen the in-tree library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdyptError(-2, f"{LIB_PATH} not built: run `python -m adypt_b200.build` (there is no CPU fallback)")
    l = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    l.adypt_last_error.restype = C.c_char_p
    sig = {
        "adypt_device_count": [vp],
        "adypt_scene_create": [C.POINTER(SceneDesc), vp],
        "adypt_scene_destroy": [vp],
        "adypt_scene_set_textures": [vp, vp, C.c_uint32],
        "adypt_scene_read_woop": [vp, vp],
        "adypt_scene_device_bytes": [vp, vp],
        "adypt_trace_closest": [vp, vp, u64, vp, vp, vp, C.c_int, vp],
        "adypt_trace_any": [vp, vp, u64, vp, C.c_int, vp],
        "adypt_launch_count": [vp],
        "adypt_trace_stats": [vp, vp, u64, C.c_int, vp],
        "adypt_trace_configure": [vp, C.c_int, C.c_int, C.c_int],
        "adypt_tracer_create": [vp, C.POINTER(PTConfig), i32, i32, u64, vp],
        "adypt_tracer_destroy": [vp],
        "adypt_tracer_set_config": [vp, C.POINTER(PTConfig)],
        "adypt_tracer_set_bias": [vp, vp],
        "adypt_tracer_get_bias": [vp, vp],
        "adypt_tracer_set_sun_visibility": [vp, i32, vp],
        "adypt_tracer_set_russian_roulette": [vp, i32],
        "adypt_tracer_set_camera": [vp, vp, vp, vp],
        "adypt_camera_matrices": [C.c_float, C.c_float, C.c_float, i32, i32, vp, vp],
        "adypt_tracer_primary": [vp, i32],
        "adypt_tracer_sample": [vp, i32],
        "adypt_tracer_accumulate": [vp, i32, i32],
        "adypt_tracer_sum_buffer": [vp, vp, vp],
        "adypt_tracer_clear_sum": [vp],
        "adypt_tracer_resolve_sum": [vp],
        "adypt_tracer_spp": [vp, vp],
        "adypt_tracer_read": [vp, vp, i32],
        "adypt_tracer_result_buffer": [vp, vp, vp],
        "adypt_tracer_save_exr": [vp, C.c_char_p, i32],
        "adypt_tracer_sync": [vp],
        "adypt_tracer_stream": [vp, vp],
        "adypt_tracer_primary_rays": [vp, vp, C.c_int],
        "adypt_tracer_stats": [vp, vp, vp],
        "adypt_sobol_vector": [C.c_uint32, C.c_uint32, vp],
        "adypt_tracer_set_profiling": [vp, i32],
        "adypt_tracer_get_profile": [vp, vp, i32],
        "adypt_trace_kernel_name": [vp, i32, C.c_char_p, u64],
        "adypt_write_exr": [C.c_char_p, vp, i32, i32, i32],
        "adypt_debug_math": [i32, i32, vp, vp, u64, vp, vp],
    }
    for name, args in sig.items():
        f = getattr(l, name)
        f.argtypes = args
        f.restype = C.c_int
    _lib = l
    return l


def _check(rc):
    if rc != 0:
        raise AdyptError(rc, load_library().adypt_last_error().decode(errors="replace"))


def _ptr(a):
    """host numpy array, torch tensor (host or device) or None -> integer address"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def _is_device(a):
    return hasattr(a, "is_cuda") and bool(a.is_cuda)


def device_count() -> int:
    n = C.c_int(0)
    _check(load_library().adypt_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    n = C.c_uint64(0)
    _check(load_library().adypt_launch_count(C.byref(n)))
    return n.value


def camera_matrices(fov, yaw, pitch, width, height):
    """Camera::GetProjection / GetView (src/Tracer/Camera.cpp:13-23) -> (proj16, view16) column-major."""
    p = np.zeros(16, dtype=np.float32)
    v = np.zeros(16, dtype=np.float32)
    _check(load_library().adypt_camera_matrices(fov, yaw, pitch, width, height, p.ctypes.data, v.ctypes.data))
    return p, v


def sobol_vector(dim: int, index: int) -> np.ndarray:
    """The vector Sobol::Next writes on call number index (0-based) after Reset(dim) (adypt_sobol_vector; host only)."""
    out = np.zeros(dim, dtype=np.float32)
    _check(load_library().adypt_sobol_vector(dim, index, out.ctypes.data))
    return out


def write_exr(path, rgb, fp16=False):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, c = rgb.shape
    assert c == 3
    _check(load_library().adypt_write_exr(path.encode(), rgb.ctypes.data, w, h, int(fp16)))


def debug_sincos(x, device=0):
    """GPU evaluation of the shading stage's deterministic sin/cos (adypt_debug_math op 0)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    _check(load_library().adypt_debug_math(device, 0, x.ctypes.data, None, x.size, s.ctypes.data, c.ctypes.data))
    return s, c


def debug_pow(x, y, device=0):
    x, y = np.ascontiguousarray(x, dtype=np.float32), np.ascontiguousarray(y, dtype=np.float32)
    o = np.empty_like(x)
    _check(load_library().adypt_debug_math(device, 1, x.ctypes.data, y.ctypes.data, x.size, o.ctypes.data, None))
    return o


class Scene:
    """Device-resident CWBVH scene: OglScene::Initialize(scene, wbvh) (OglScene.cpp:46-49, 118-141)."""

    def __init__(self, nodes, tri_indices, woop=None, triangles=None, materials=None, device=0):
        lib = load_library()
        nodes = np.ascontiguousarray(nodes, dtype=np.uint8).reshape(-1, 80)
        tri_indices = np.ascontiguousarray(tri_indices, dtype=np.int32)
        self.n_nodes, self.n_refs = nodes.shape[0], tri_indices.shape[0]
        d = SceneDesc()
        d.device = device
        d.nodes, d.n_nodes = nodes.ctypes.data, self.n_nodes
        d.tri_indices, d.n_refs = tri_indices.ctypes.data, self.n_refs
        keep = [nodes, tri_indices]
        if woop is not None:
            woop = np.ascontiguousarray(woop, dtype=np.float32).reshape(-1, 12)
            assert woop.shape[0] == self.n_refs
            d.woop = woop.ctypes.data
            keep.append(woop)
        if triangles is not None:
            triangles = np.ascontiguousarray(triangles, dtype=np.uint8).reshape(-1, 100)
            d.triangles, d.n_tris = triangles.ctypes.data, triangles.shape[0]
            keep.append(triangles)
        if materials is not None:
            materials = np.ascontiguousarray(materials, dtype=np.uint8).reshape(-1, 64)
            d.materials, d.n_mats = materials.ctypes.data, materials.shape[0]
            keep.append(materials)
        self.device = device
        self._h = C.c_void_p()
        _check(lib.adypt_scene_create(C.byref(d), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            try:
                load_library().adypt_scene_destroy(self._h)
            except Exception:  # interpreter shutdown: module globals may already be gone
                pass
            self._h = None

    __del__ = close

    def set_textures(self, textures):
        """Diffuse textures (list of (h,w,3) uint8 arrays); texture i is what material.dtex == i refers to."""
        arrs = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
        descs = (TextureDesc * max(1, len(arrs)))()
        for i, a in enumerate(arrs):
            assert a.ndim == 3 and a.shape[2] == 3
            descs[i] = TextureDesc(a.ctypes.data, a.shape[1], a.shape[0])
        _check(load_library().adypt_scene_set_textures(self._h, descs, len(arrs)))

    def read_woop(self):
        out = np.zeros((self.n_refs, 12), dtype=np.float32)
        _check(load_library().adypt_scene_read_woop(self._h, out.ctypes.data))
        return out

    def device_bytes(self):
        b = C.c_uint64(0)
        _check(load_library().adypt_scene_device_bytes(self._h, C.byref(b)))
        return b.value

    def configure(self, ctas_per_sm=0, refill_threshold=0, variant=0):
        _check(load_library().adypt_trace_configure(self._h, ctas_per_sm, refill_threshold, variant))

    def trace_closest(self, rays, tri=None, t=None, uv=None, stream=None, want_t=True, want_uv=True):
        """Batch closest hit. `rays`: (n,8) float32 numpy array (host) or CUDA torch tensor (device).
        Host call: returns dict(tri, t, uv) numpy arrays. Device call: outputs must be CUDA tensors
        (tri int32 (n,), t float32 (n,) or None, uv float32 (n,2) or None); asynchronous on `stream`."""
        lib = load_library()
        if _is_device(rays):
            n = rays.numel() // 8
            assert tri is not None
            _check(lib.adypt_trace_closest(self._h, _ptr(rays), n, _ptr(tri), _ptr(t), _ptr(uv), MEM_DEVICE, stream))
            return dict(tri=tri, t=t, uv=uv)
        if hasattr(rays, "numpy") and not isinstance(rays, np.ndarray):  # pinned host torch tensor
            n = rays.numel() // 8
            _check(lib.adypt_trace_closest(self._h, _ptr(rays), n, _ptr(tri), _ptr(t), _ptr(uv), MEM_HOST, stream))
            return dict(tri=tri, t=t, uv=uv)
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.size // 8
        tri = np.empty(n, dtype=np.int32) if tri is None else tri
        t = (np.empty(n, dtype=np.float32) if want_t else None) if t is None else t
        uv = (np.zeros((n, 2), dtype=np.float32) if want_uv else None) if uv is None else uv
        _check(lib.adypt_trace_closest(self._h, rays.ctypes.data, n, _ptr(tri), _ptr(t), _ptr(uv), MEM_HOST, stream))
        return dict(tri=tri, t=t, uv=uv)

    def trace_stats(self, rays):
        """Instrumented pass: dict(nodes, tris, hits, max_stack) totals over the batch (adypt_trace_stats)."""
        out = (C.c_uint64 * 4)()
        if _is_device(rays):
            _check(load_library().adypt_trace_stats(self._h, _ptr(rays), rays.numel() // 8, MEM_DEVICE, out))
        else:
            rays = np.ascontiguousarray(rays, dtype=np.float32)
            _check(load_library().adypt_trace_stats(self._h, rays.ctypes.data, rays.size // 8, MEM_HOST, out))
        return dict(nodes=int(out[0]), tris=int(out[1]), hits=int(out[2]), max_stack=int(out[3]))

    def kernel_name(self, any_hit=False, demangle=True):
        """Name of the kernel trace_closest / trace_any launch for the current tuning variant (adypt_trace_kernel_name)."""
        buf = C.create_string_buffer(512)
        _check(load_library().adypt_trace_kernel_name(self._h, 1 if any_hit else 0, buf, 512))
        name = buf.value.decode()
        if demangle:
            import shutil
            import subprocess
            tool = shutil.which("c++filt") or shutil.which("cu++filt")
            if tool:
                try:
                    name = subprocess.run([tool, name], capture_output=True, text=True, timeout=10).stdout.strip() or name
                except Exception:
                    pass
        return name

    def trace_any(self, rays, occluded=None, stream=None):
        lib = load_library()
        if _is_device(rays):
            n = rays.numel() // 8
            assert occluded is not None
            _check(lib.adypt_trace_any(self._h, _ptr(rays), n, _ptr(occluded), MEM_DEVICE, stream))
            return occluded
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.size // 8
        occluded = np.empty(n, dtype=np.uint8) if occluded is None else occluded
        _check(lib.adypt_trace_any(self._h, rays.ctypes.data, n, _ptr(occluded), MEM_HOST, stream))
        return occluded


class Tracer:
    """OglPathTracer (src/Tracer/OglPathTracer.hpp:66-82) over the wavefront CUDA integrator."""

    def __init__(self, scene: Scene, config: PTConfig, width: int, height: int, bias_seed: int = 0):
        self.scene, self.width, self.height = scene, width, height
        self.config = config
        self._h = C.c_void_p()
        _check(load_library().adypt_tracer_create(scene._h, C.byref(config), width, height, bias_seed, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            try:
                load_library().adypt_tracer_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close

    def set_config(self, config: PTConfig):
        self.config = config
        _check(load_library().adypt_tracer_set_config(self._h, C.byref(config)))

    def set_bias(self, rg8):
        rg8 = np.ascontiguousarray(rg8, dtype=np.uint8)
        assert rg8.size == self.width * self.height * 2
        _check(load_library().adypt_tracer_set_bias(self._h, rg8.ctypes.data))

    def get_bias(self):
        out = np.zeros((self.height, self.width, 2), dtype=np.uint8)
        _check(load_library().adypt_tracer_get_bias(self._h, out.ctypes.data))
        return out

    def set_russian_roulette(self, start_bounce):
        """Opt-in extension (not in the reference): roulette from bounce `start_bounce` on; None / negative = off."""
        _check(load_library().adypt_tracer_set_russian_roulette(self._h, -1 if start_bounce is None else int(start_bounce)))

    def set_sun_visibility(self, enabled: bool, direction=(0.6, 1.0, 0.2)):
        """Connect stage: the any-hit sun test the reference has commented out (pathtracer.glsl:132)."""
        d = np.ascontiguousarray(direction, dtype=np.float32)
        _check(load_library().adypt_tracer_set_sun_visibility(self._h, int(enabled), d.ctypes.data))

    def set_camera(self, projection, view, position):
        """SetCamera(projection, view, position) (OglPathTracer.cpp:27-32)."""
        p = np.ascontiguousarray(projection, dtype=np.float32).reshape(16)
        v = np.ascontiguousarray(view, dtype=np.float32).reshape(16)
        o = np.ascontiguousarray(position, dtype=np.float32).reshape(3)
        _check(load_library().adypt_tracer_set_camera(self._h, p.ctypes.data, v.ctypes.data, o.ctypes.data))

    def look(self, position, yaw, pitch, fov):
        """Camera::GetProjection/GetView + SetCamera, as Instance::Update does (Instance.cpp:50-52)."""
        p, v = camera_matrices(fov, yaw, pitch, self.width, self.height)
        self.set_camera(p, v, position)
        return p, v

    def trace(self, enable_pt: bool, viewer_type: int = VIEW_DIFFUSE, n_spp: int = 1):
        """Trace(bool) (OglPathTracer.cpp:34-61)."""
        if enable_pt:
            _check(load_library().adypt_tracer_sample(self._h, n_spp))
        else:
            _check(load_library().adypt_tracer_primary(self._h, viewer_type))

    def sample(self, n_spp: int):
        _check(load_library().adypt_tracer_sample(self._h, n_spp))

    def primary(self, viewer_type: int):
        _check(load_library().adypt_tracer_primary(self._h, viewer_type))

    def accumulate(self, first_spp: int, n_spp: int):
        _check(load_library().adypt_tracer_accumulate(self._h, first_spp, n_spp))

    def clear_sum(self):
        _check(load_library().adypt_tracer_clear_sum(self._h))

    def resolve_sum(self):
        _check(load_library().adypt_tracer_resolve_sum(self._h))

    def sum_buffer(self):
        p, n = C.c_void_p(), C.c_uint64()
        _check(load_library().adypt_tracer_sum_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def result_buffer(self):
        p, n = C.c_void_p(), C.c_uint64()
        _check(load_library().adypt_tracer_result_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    @property
    def spp(self) -> int:
        v = C.c_int32(0)
        _check(load_library().adypt_tracer_spp(self._h, C.byref(v)))
        return v.value

    def stream(self) -> int:
        """The cudaStream_t (as an integer) the tracer enqueues its work on (adypt_tracer_stream)."""
        p = C.c_void_p()
        _check(load_library().adypt_tracer_stream(self._h, C.byref(p)))
        return p.value or 0

    def sync(self):
        _check(load_library().adypt_tracer_sync(self._h))

    def read(self, channels=4, out=None):
        if out is None:
            out = np.empty((self.height, self.width, channels), dtype=np.float32)
        _check(load_library().adypt_tracer_read(self._h, _ptr(out), channels))
        return out

    def save_exr(self, path: str, fp16: bool = False):
        """SaveResult(filename, save_as_fp16) (OglPathTracer.cpp:199-212)."""
        _check(load_library().adypt_tracer_save_exr(self._h, path.encode(), int(fp16)))

    def primary_rays(self, out=None):
        """Pinhole rays of primaryray.glsl:39-44 for the current camera; (h*w, 8) float32."""
        if out is not None and _is_device(out):
            _check(load_library().adypt_tracer_primary_rays(self._h, _ptr(out), MEM_DEVICE))
            return out
        out = np.empty((self.width * self.height, 8), dtype=np.float32)
        _check(load_library().adypt_tracer_primary_rays(self._h, out.ctypes.data, MEM_HOST))
        return out

    def set_profiling(self, stage_times=False, trace_counters=False):
        """Measurement hooks (adypt_tracer_set_profiling): event pairs around every stage / the instrumented traversal kernel."""
        _check(load_library().adypt_tracer_set_profiling(self._h, (1 if stage_times else 0) | (2 if trace_counters else 0)))

    def profile(self, reset=True):
        """dict(stage_ms={name: ms}, stage_launches={name: n}, trace=dict(...) for the bounce queues, primary=dict(...) for the primary rays) since the last reset."""
        p = TracerProfile()
        _check(load_library().adypt_tracer_get_profile(self._h, C.byref(p), 1 if reset else 0))
        return dict(stage_ms={n: p.stage_ms[i] for i, n in enumerate(STAGES)}, stage_launches={n: int(p.stage_launches[i]) for i, n in enumerate(STAGES)},
                    trace=dict(nodes=int(p.trace_nodes), tris=int(p.trace_tris), hits=int(p.trace_hits), rays=int(p.trace_rays), max_stack=int(p.trace_max_depth)),
                    primary=dict(nodes=int(p.primary_nodes), tris=int(p.primary_tris), hits=int(p.primary_hits), rays=int(p.primary_rays)))

    def stats(self):
        s, l = C.c_uint64(0), C.c_uint64(0)
        _check(load_library().adypt_tracer_stats(self._h, C.byref(s), C.byref(l)))
        return dict(segments=s.value, launches=l.value)
