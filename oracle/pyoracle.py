"""TEST INFRASTRUCTURE ONLY.  ctypes bindings for the oracle libraries:

* ``oracle/_ref/libvrm_ref_host.so``  -- the UNMODIFIED reference headers built for the host (prefix ``refh_``),
* ``oracle/liboracle_port.so``        -- the plain-C restatement in oracle/vrm_oracle.c (prefix ``orc_``),
* ``oracle/_ref/libvrm_ref_cuda*.so`` -- the reference's own CUDA kernels rebuilt for sm_100a (prefix ``refg_``).

Only tests/, ``__graft_entry__.smoke()`` and the baseline legs of bench.py may import this module.
The three libraries expose the same C entry points, so one wrapper class serves all of them.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBS = {
    "refh": os.path.join(HERE, "_ref", "libvrm_ref_host.so"),
    "orc": os.path.join(HERE, "liboracle_port.so"),
    "refg": os.path.join(HERE, "_ref", "libvrm_ref_cuda.so"),
    "refgx": os.path.join(HERE, "_ref", "libvrm_ref_cuda_exact.so"),
    # host compile of the PRODUCT's traversal core (tests/hostsim/hostsim.cpp) -- a device-code debugging aid, not an oracle
    "sim": os.path.join(HERE, "..", "tests", "hostsim", "libhostsim.so"),
}
PREFIX = {"refh": "refh", "orc": "orc", "refg": "refg", "refgx": "refg", "sim": "sim"}

STORAGE = {"vcs": 0, "hashtable": 1}          # StorageType, VoxelFunctions.cuh:37
ALGORITHM = {"longestaxis": 0, "original": 1}  # Main.cu:58-68
EMPTY = 1 << 30

_loaded = {}


def available(kind: str) -> bool:
    return os.path.exists(LIBS[kind])


def _lib(kind: str):
    if kind not in _loaded:
        lib = C.CDLL(LIBS[kind])
        p = PREFIX[kind]
        vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_float
        sig = {
            "scene_create": (vp, []),
            "scene_destroy": (None, [vp]),
            "scene_add_voxels": (None, [vp, vp, vp, u64]),
            "scene_build": (C.c_int, [vp, C.c_int]),
            "scene_info": (None, [vp, vp, vp, vp]),
            "set_lighting": (None, [vp, vp, vp, C.c_int, C.c_int]),
            "make_unit_vector": (None, [vp, vp]),
            "camera_make": (None, [vp, vp, vp, f32, f32, vp]),
            "render": (C.c_int, [vp, vp, vp, u32, C.c_int, u32, u32, vp, vp, vp, vp, C.c_int]),
            "trace_rays": (C.c_int, [vp, vp, u64, vp, u32, C.c_int, vp, vp, vp, C.c_int]),
            "lookup": (C.c_int, [vp, vp, u64, vp, vp]),
        }
        if kind == "refh":
            sig["scene_add_cube"] = (None, [vp, i32, i32, i32, i32])
            sig["scene_add_sphere"] = (None, [vp, u32, u32, u32, u32, C.c_int])
            sig["scene_load_file"] = (None, [vp, C.c_char_p])
        if kind in ("refg", "refgx"):
            sig["last_kernel_ms"] = (C.c_float, [])
            sig["render_timed"] = (C.c_int, [vp, vp, vp, u32, C.c_int, u32, u32, C.c_int, C.c_int, vp])
        for name, (res, args) in sig.items():
            fn = getattr(lib, f"{p}_{name}")
            fn.restype = res
            fn.argtypes = args
        _loaded[kind] = lib
    return _loaded[kind]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, np.float32).reshape(3))


def unit_vector(v, kind="refh"):
    out = np.zeros(3, np.float32)
    getattr(_lib(kind), f"{PREFIX[kind]}_make_unit_vector")(_ptr(_f3(v)), _ptr(out))
    return out


def make_camera(origin, look_at, up, fov, aspect, kind="refh"):
    """Camera.cuh:11-23 -> the 60-byte struct as 15 floats."""
    out = np.zeros(15, np.float32)
    getattr(_lib(kind), f"{PREFIX[kind]}_camera_make")(_ptr(_f3(origin)), _ptr(_f3(look_at)), _ptr(_f3(up)),
                                                       C.c_float(np.float32(fov)), C.c_float(np.float32(aspect)), _ptr(out))
    return out


DEFAULT_LIGHT = dict(direction=None, colour=(1.0, 1.0, 1.0), position=(10.0, 10.0, -10.0), use_point=False, use_shadows=True)


def set_lighting(kind="refh", direction=None, colour=(1.0, 1.0, 1.0), position=(10.0, 10.0, -10.0), use_point=False, use_shadows=True):
    """Main.cu:26-42 (process-wide constants, as in the reference).  direction None = unit(1,1,1)."""
    if direction is None:
        direction = unit_vector((1.0, 1.0, 1.0), kind)
    getattr(_lib(kind), f"{PREFIX[kind]}_set_lighting")(_ptr(_f3(direction)), _ptr(_f3(colour)), _ptr(_f3(position)),
                                                        int(use_point), int(use_shadows))


class OracleScene:
    def __init__(self, kind: str = "refh"):
        self.kind = kind
        self.lib = _lib(kind)
        self.p = PREFIX[kind]
        self.h = self._fn("scene_create")()
        self.storage = None

    def _fn(self, name):
        return getattr(self.lib, f"{self.p}_{name}")

    def close(self):
        if self.h:
            self._fn("scene_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_voxels(self, xyz, rgb):
        xyz = np.ascontiguousarray(xyz, np.int32).reshape(-1, 3)
        rgb = np.ascontiguousarray(rgb, np.uint32).reshape(-1)
        assert xyz.shape[0] == rgb.shape[0]
        self._fn("scene_add_voxels")(self.h, _ptr(xyz), _ptr(rgb), xyz.shape[0])

    def load_file(self, directory: str, filename: str = "scene.vox"):
        """VoxelFile::readVoxelFile on ``directory``/resources/``filename`` (the reference resolves the path against the working directory)."""
        here = os.getcwd()
        os.chdir(directory)
        try:
            self._fn("scene_load_file")(self.h, filename.encode())
        finally:
            os.chdir(here)

    def add_cube(self, x, y, z, hw):
        self._fn("scene_add_cube")(self.h, x, y, z, hw)

    def add_sphere(self, x, y, z, r, checkered=False):
        self._fn("scene_add_sphere")(self.h, x, y, z, r, int(checkered))

    def build(self, storage: str):
        rc = self._fn("scene_build")(self.h, STORAGE[storage])
        if rc:
            raise RuntimeError(f"{self.p}_scene_build failed: {rc}")
        self.storage = storage

    def info(self):
        d, m, f = C.c_uint32(), C.c_int32(), C.c_uint32()
        self._fn("scene_info")(self.h, C.byref(d), C.byref(m), C.byref(f))
        return dict(diameter=d.value, min_coord=m.value, filled=f.value)

    def render(self, camera15, width, height, algorithm: str, scale=1, translation=(0, 0, 0), want_hits=True,
               want_counters=False, want_lookups=False, threads=None):
        rgb = np.zeros((height, width, 3), np.uint8)
        hits = np.zeros((height, width, 4), np.int32) if want_hits else None
        counters = np.zeros(5, np.uint64) if want_counters else None
        lookups = np.zeros((height, width), np.uint32) if want_lookups else None
        cam = np.ascontiguousarray(camera15, np.float32)
        rc = self._fn("render")(self.h, _ptr(cam), _ptr(_f3(translation)), scale, ALGORITHM[algorithm], width, height,
                                _ptr(rgb), _ptr(hits), _ptr(counters), _ptr(lookups), threads or (os.cpu_count() or 1))
        if rc:
            raise RuntimeError(f"{self.p}_render failed: {rc}")
        return dict(rgb=rgb, hits=hits, counters=counters, lookups=lookups)

    def trace_rays(self, rays, algorithm: str, scale=1, translation=(0, 0, 0), want_hits=True, want_counters=False, threads=None):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        colour = np.zeros(n, np.uint32)
        hits = np.zeros((n, 4), np.int32) if want_hits else None
        counters = np.zeros(5, np.uint64) if want_counters else None
        rc = self._fn("trace_rays")(self.h, _ptr(rays), n, _ptr(_f3(translation)), scale, ALGORITHM[algorithm],
                                    _ptr(colour), _ptr(hits), _ptr(counters), threads or (os.cpu_count() or 1))
        if rc:
            raise RuntimeError(f"{self.p}_trace_rays failed: {rc}")
        return dict(colour=colour, hits=hits, counters=counters)

    def lookup(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.int32).reshape(-1, 3)
        n = xyz.shape[0]
        out = np.zeros(n, np.uint32)
        exists = np.zeros(n, np.uint8)
        rc = self._fn("lookup")(self.h, _ptr(xyz), n, _ptr(out), _ptr(exists))
        if rc:
            raise RuntimeError(f"{self.p}_lookup failed: {rc}")
        return out, exists
