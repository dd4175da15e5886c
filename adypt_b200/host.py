"""Host-side stages that feed the tracer (no GPU needed): Triangle[] assembly, the from-scratch SBVH -> CWBVH
builder (byte-identical to the reference's src/BVH pipeline), the .bvh cache file and OBJ/MTL ingest.
Thin ctypes binding over the "Host side" section of include/adypt_b200.h."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import AdyptError, Scene, _check, load_library  # noqa: F401


class BvhConfig(C.Structure):
    """InstanceConfig::BVH (src/InstanceConfig.hpp:15-20)."""
    _fields_ = [("max_spatial_depth", C.c_int32), ("triangle_sah", C.c_float), ("node_sah", C.c_float)]

    @classmethod
    def make(cls, max_spatial_depth=48, triangle_sah=0.3, node_sah=1.0):
        return cls(max_spatial_depth, triangle_sah, node_sah)


class CamConfig(C.Structure):
    """InstanceConfig::Cam (src/InstanceConfig.hpp:30-35)."""
    _fields_ = [("speed", C.c_float), ("mouse_sensitive", C.c_float), ("fov", C.c_float), ("yaw", C.c_float),
                ("pitch", C.c_float), ("position", C.c_float * 3)]


class InstanceConfig(C.Structure):
    """InstanceConfig (src/InstanceConfig.hpp:12-48): the .config file."""
    from . import PTConfig as _PT
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bvh", BvhConfig), ("pt", _PT), ("cam", CamConfig),
                ("obj_filename", C.c_char * 1024), ("bvh_filename", C.c_char * 1024)]

    @classmethod
    def default(cls):
        c = cls()
        _check(_lib().adypt_config_set_default(C.byref(c)))
        return c

    @classmethod
    def load(cls, path: str):
        """InstanceConfig::LoadFromFile; raises AdyptError carrying the reference's [PARSER]ERR text."""
        c = cls()
        _check(_lib().adypt_config_load(path.encode(), C.byref(c)))
        return c

    def to_json(self) -> str:
        need = C.c_uint64(0)
        _check(_lib().adypt_config_to_json(C.byref(self), None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(_lib().adypt_config_to_json(C.byref(self), buf, need.value, None))
        return buf.value.decode()

    def save(self, path: str):
        _check(_lib().adypt_config_save(C.byref(self), path.encode()))


class _Info(C.Structure):
    _fields_ = [("n_tris", C.c_uint32), ("n_mats", C.c_uint32), ("n_nodes", C.c_uint32), ("n_refs", C.c_uint32),
                ("n_binary_nodes", C.c_uint32), ("triangles", C.c_void_p), ("materials", C.c_void_p), ("nodes", C.c_void_p),
                ("tri_indices", C.c_void_p), ("binary_nodes", C.c_void_p), ("aabb", C.c_float * 6)]


_bound = False


def _lib():
    global _bound
    l = load_library()
    if not _bound:
        vp = C.c_void_p
        l.adypt_host_scene_load_obj.argtypes = [C.c_char_p, vp]
        l.adypt_host_scene_from_triangles.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp]
        l.adypt_host_scene_destroy.argtypes = [vp]
        l.adypt_host_scene_build_bvh.argtypes = [vp, C.POINTER(BvhConfig)]
        l.adypt_host_scene_load_bvh.argtypes = [vp, C.c_char_p, C.POINTER(BvhConfig)]
        l.adypt_host_scene_save_bvh.argtypes = [vp, C.c_char_p, C.POINTER(BvhConfig)]
        l.adypt_host_scene_get.argtypes = [vp, C.POINTER(_Info)]
        l.adypt_host_scene_upload.argtypes = [vp, C.c_int32, vp]
        for n in ("load_obj", "from_triangles", "destroy", "build_bvh", "load_bvh", "save_bvh", "get", "upload"):
            getattr(l, "adypt_host_scene_" + n).restype = C.c_int
        l.adypt_host_scene_load_textures.argtypes = [vp, vp, vp]
        l.adypt_host_scene_load_textures.restype = C.c_int
        l.adypt_host_scene_texture.argtypes = [vp, C.c_uint32, vp, vp, vp]
        l.adypt_host_scene_texture.restype = C.c_int
        l.adypt_config_set_default.argtypes = [vp]
        l.adypt_config_load.argtypes = [C.c_char_p, vp]
        l.adypt_config_to_json.argtypes = [vp, vp, C.c_uint64, vp]
        l.adypt_config_save.argtypes = [vp, C.c_char_p]
        _bound = True
    return l


def _copy(ptr, nbytes, dtype):
    if not ptr or nbytes == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_char * nbytes).from_address(ptr), dtype=dtype).copy()


def materials_array(materials) -> np.ndarray:
    """workloads.Material list -> (n,64) uint8 GPUMaterial records (OglScene.hpp:19-28)."""
    out = np.zeros((len(materials), 16), dtype=np.float32)
    iv = out.view(np.int32)
    for i, m in enumerate(materials):
        iv[i, 0] = -1
        out[i, 1:4] = m.kd
        out[i, 5:8] = m.ke
        out[i, 9:12] = m.ks
        iv[i, 12] = m.illum
        out[i, 13], out[i, 14], out[i, 15] = m.ns, m.d, m.ni
    return out.view(np.uint8).reshape(len(materials), 64)


class HostScene:
    """Scene (src/Util/Scene.hpp) + WideBVH (src/BVH/WideBVH.hpp) on the host."""

    def __init__(self, handle):
        self._h = handle
        self._refresh()

    @classmethod
    def from_obj(cls, path: str):
        h = C.c_void_p()
        _check(_lib().adypt_host_scene_load_obj(path.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def from_triangles(cls, positions, material_ids, materials64):
        p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 9)
        m = np.ascontiguousarray(material_ids, dtype=np.int32)
        mats = np.ascontiguousarray(materials64, dtype=np.uint8).reshape(-1, 64)
        h = C.c_void_p()
        _check(_lib().adypt_host_scene_from_triangles(p.ctypes.data, m.ctypes.data, p.shape[0], mats.ctypes.data, mats.shape[0], C.byref(h)))
        return cls(h)

    def close(self):
        if getattr(self, "_h", None):
            try:
                _lib().adypt_host_scene_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close

    def _refresh(self):
        i = _Info()
        _check(_lib().adypt_host_scene_get(self._h, C.byref(i)))
        self.tris = _copy(i.triangles, i.n_tris * 100, np.uint8).reshape(-1, 100)
        self.mats = _copy(i.materials, i.n_mats * 64, np.uint8).reshape(-1, 64)
        self.nodes = _copy(i.nodes, i.n_nodes * 80, np.uint8).reshape(-1, 80)
        self.tri_indices = _copy(i.tri_indices, i.n_refs * 4, np.int32)
        self.binary_nodes = _copy(i.binary_nodes, i.n_binary_nodes * 32, np.uint8).reshape(-1, 32)
        self.aabb = np.array(list(i.aabb), dtype=np.float32)

    def build_bvh(self, config: BvhConfig | None = None):
        cfg = config or BvhConfig.make()
        _check(_lib().adypt_host_scene_build_bvh(self._h, C.byref(cfg)))
        self._refresh()
        return self

    def load_bvh(self, path: str, config: BvhConfig | None = None) -> bool:
        cfg = config or BvhConfig.make()
        rc = _lib().adypt_host_scene_load_bvh(self._h, path.encode(), C.byref(cfg))
        if rc == -5:
            return False
        _check(rc)
        self._refresh()
        return True

    def save_bvh(self, path: str, config: BvhConfig | None = None):
        cfg = config or BvhConfig.make()
        _check(_lib().adypt_host_scene_save_bvh(self._h, path.encode(), C.byref(cfg)))

    def positions(self):
        return self.tris[:, :36].copy().view(np.float32).reshape(-1, 3, 3)

    def load_textures(self):
        """Decode the materials' map_Kd files and renumber Material::dtex like OglScene::init_materials.
        Returns (n_loaded, n_failed); the decoded images are then in .textures ((h,w,3) uint8 arrays)."""
        ok, bad = C.c_uint32(0), C.c_uint32(0)
        _check(_lib().adypt_host_scene_load_textures(self._h, C.byref(ok), C.byref(bad)))
        self.textures = []
        for i in range(ok.value):
            p, w, h = C.c_void_p(), C.c_int32(), C.c_int32()
            _check(_lib().adypt_host_scene_texture(self._h, i, C.byref(p), C.byref(w), C.byref(h)))
            self.textures.append(_copy(p.value, w.value * h.value * 3, np.uint8).reshape(h.value, w.value, 3))
        self._refresh()
        return ok.value, bad.value

    def upload(self, device: int = 0) -> Scene:
        """OglScene::Initialize(scene, wbvh) (Instance.cpp:33)."""
        h = C.c_void_p()
        _check(_lib().adypt_host_scene_upload(self._h, device, C.byref(h)))
        s = Scene.__new__(Scene)
        s._h, s.device, s.n_nodes, s.n_refs = h, device, self.nodes.shape[0], self.tri_indices.shape[0]
        return s


class RenderGroup:
    """adypt_group_*: one process driving several GPUs (sample-sharded render + one NCCL reduce)."""

    def __init__(self, host_scene: HostScene, config, width: int, height: int, devices, bias_seed: int = 0):
        l = _lib()
        vp = C.c_void_p
        l.adypt_group_create.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_uint64, vp, C.c_uint32, vp]
        l.adypt_group_destroy.argtypes = [vp]
        l.adypt_group_set_camera.argtypes = [vp, vp, vp, vp]
        l.adypt_group_set_sun_visibility.argtypes = [vp, C.c_int32, vp]
        l.adypt_group_set_russian_roulette.argtypes = [vp, C.c_int32]
        l.adypt_group_render.argtypes = [vp, C.c_int32]
        l.adypt_group_read.argtypes = [vp, vp, C.c_int32]
        l.adypt_group_save_exr.argtypes = [vp, C.c_char_p, C.c_int32]
        self.width, self.height = width, height
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        self._h = C.c_void_p()
        _check(l.adypt_group_create(host_scene._h, C.byref(config), width, height, bias_seed, devs.ctypes.data, devs.size, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            try:
                _lib().adypt_group_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close

    def look(self, position, yaw, pitch, fov):
        from . import camera_matrices
        p, v = camera_matrices(fov, yaw, pitch, self.width, self.height)
        o = np.ascontiguousarray(position, dtype=np.float32)
        _check(_lib().adypt_group_set_camera(self._h, p.ctypes.data, v.ctypes.data, o.ctypes.data))

    def render(self, total_spp: int):
        _check(_lib().adypt_group_render(self._h, total_spp))

    def read(self, channels=4):
        out = np.empty((self.height, self.width, channels), dtype=np.float32)
        _check(_lib().adypt_group_read(self._h, out.ctypes.data, channels))
        return out

    def save_exr(self, path: str, fp16: bool = False):
        _check(_lib().adypt_group_save_exr(self._h, path.encode(), int(fp16)))


def build_scene(mesh, config: BvhConfig | None = None) -> HostScene:
    """workloads.SceneMesh -> HostScene with its CWBVH built (Instance.cpp:12-24 without the OBJ detour)."""
    hs = HostScene.from_triangles(mesh.positions(), mesh.face_mat, materials_array(mesh.materials))
    return hs.build_bvh(config)
