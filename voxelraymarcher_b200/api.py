"""Host-side mirror of the reference's interface for the hot path, over the C ABI of ``libvrm_b200.so``
(``include/vrm_b200.h``).  Names follow the reference (VoxelRaymarcher/src):

* ``VoxelScene.insert_voxel / add_voxels``  <- ``VoxelSceneCPU::insertVoxel``            geometry/VoxelSceneCPU.cuh:16-46
* ``VoxelScene.generate_voxel_scene``       <- ``VoxelSceneCPU::generateVoxelScene``     geometry/VoxelSceneCPU.cuh:49-93
* ``VoxelScene.get_array_diameter / get_array_size / get_min_coord``                    geometry/VoxelSceneCPU.cuh:107-123
* ``Camera``                                <- ``Camera::Camera``                        renderer/camera/Camera.cuh:11-23
* ``VoxelScene.setup_constant_values``      <- ``setupConstantValues``                   main/Main.cu:26-42
* ``VoxelScene.run_raymarching_kernel``     <- ``runRaymarchingKernel``                  main/Main.cu:105-163
* ``StorageType`` / algorithm ids                                                       geometry/VoxelFunctions.cuh:37, main/Main.cu:58-68

There is NO CPU fallback: importing works anywhere (so CPU-only tests can check the library's exports), but every
compute entry point raises ``VrmError`` when the CUDA library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VRM_B200_LIB") or os.path.join(HERE, "libvrm_b200.so")   # VRM_B200_LIB: an A/B build of the same library (tools/ab_variants.py)

STORAGE_VCS, STORAGE_HASHTABLE = 0, 1          # StorageType {VOXEL_CLUSTER_STORE, HASH_TABLE}
ALGO_LONGEST_AXIS, ALGO_ORIGINAL = 0, 1        # rayMarchFunctionID
STORAGE = {"vcs": STORAGE_VCS, "hashtable": STORAGE_HASHTABLE}
ALGORITHM = {"longestaxis": ALGO_LONGEST_AXIS, "original": ALGO_ORIGINAL}
EMPTY = 1 << 30                                # EMPTY_KEY / EMPTY_VAL

EXPORTS = [
    "vrm_error_string", "vrm_last_error", "vrm_device_available", "vrm_device_count", "vrm_device_name", "vrm_scene_create", "vrm_scene_destroy",
    "vrm_scene_set_stream", "vrm_scene_reset_stream", "vrm_scene_synchronize", "vrm_scene_add_voxels", "vrm_scene_add_voxels_device", "vrm_scene_insert_voxel",
    "vrm_scene_generate_terrain", "vrm_scene_generate_sparse_shells", "vrm_scene_generate_cube", "vrm_scene_generate_sphere",
    "vrm_scene_build", "vrm_scene_info", "vrm_set_lighting", "vrm_camera_make", "vrm_make_unit_vector", "vrm_render",
    "vrm_render_device", "vrm_render_views_device", "vrm_render_views_device_strided", "vrm_render_views", "vrm_trace_rays", "vrm_trace_rays_device", "vrm_lookup",
    "vrm_set_l2_persistence", "vrm_set_statistics", "vrm_get_statistics", "vrm_peer_alloc", "vrm_peer_open", "vrm_peer_close", "vrm_peer_free", "vrm_copy_device", "vrm_microbench_l2",
    "vrm_scene_set_completion_flag", "vrm_scene_set_completion_counter", "vrm_claim_next", "vrm_wait_flags_device", "vrm_render_views_sharded",
]


class VrmError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libvrm_b200.so (built in-tree by ``__graft_entry__.build()`` / ``make -C voxelraymarcher_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VrmError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, ci, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int, C.c_float
    sig = {
        "vrm_error_string": (C.c_char_p, [ci]),
        "vrm_last_error": (C.c_char_p, [vp]),
        "vrm_device_available": (ci, []),
        "vrm_device_count": (ci, []),
        "vrm_device_name": (ci, [ci, C.c_char_p, u64]),
        "vrm_scene_create": (ci, [ci, C.POINTER(vp)]),
        "vrm_scene_destroy": (ci, [vp]),
        "vrm_scene_set_stream": (ci, [vp, vp]),
        "vrm_scene_reset_stream": (ci, [vp]),
        "vrm_scene_synchronize": (ci, [vp]),
        "vrm_scene_add_voxels": (ci, [vp, vp, vp, u64]),
        "vrm_scene_add_voxels_device": (ci, [vp, vp, vp, u64]),
        "vrm_scene_insert_voxel": (ci, [vp, i32, i32, i32, u32]),
        "vrm_scene_generate_terrain": (ci, [vp, u32, u32, u32, C.POINTER(u64)]),
        "vrm_scene_generate_sparse_shells": (ci, [vp, u32, u32, u32, u32, C.POINTER(u64)]),
        "vrm_scene_generate_cube": (ci, [vp, i32, i32, i32, i32, C.POINTER(u64)]),
        "vrm_scene_generate_sphere": (ci, [vp, u32, u32, u32, u32, ci, C.POINTER(u64)]),
        "vrm_scene_build": (ci, [vp, ci, C.POINTER(f32)]),
        "vrm_scene_info": (ci, [vp, C.POINTER(u32), C.POINTER(i32), C.POINTER(u32), C.POINTER(u64), C.POINTER(u64)]),
        "vrm_set_lighting": (ci, [vp, vp, vp, vp, ci, ci]),
        "vrm_camera_make": (ci, [vp, vp, vp, f32, f32, vp]),
        "vrm_make_unit_vector": (ci, [vp, vp]),
        "vrm_render": (ci, [vp, vp, vp, u32, ci, u32, u32, vp, vp, C.POINTER(f32)]),
        "vrm_render_device": (ci, [vp, vp, vp, u32, ci, u32, u32, vp, vp]),
        "vrm_render_views_device": (ci, [vp, vp, u32, vp, u32, ci, u32, u32, vp, vp]),
        "vrm_render_views_device_strided": (ci, [vp, vp, u32, vp, u32, ci, u32, u32, vp, vp, u32]),
        "vrm_render_views": (ci, [vp, vp, u32, vp, u32, ci, u32, u32, vp, C.POINTER(f32)]),
        "vrm_trace_rays": (ci, [vp, vp, u64, vp, u32, ci, vp, vp, C.POINTER(f32)]),
        "vrm_trace_rays_device": (ci, [vp, vp, u64, vp, u32, ci, vp, vp]),
        "vrm_lookup": (ci, [vp, vp, u64, vp, vp]),
        "vrm_peer_alloc": (ci, [ci, u64, C.POINTER(vp), vp]),
        "vrm_peer_open": (ci, [ci, vp, C.POINTER(vp)]),
        "vrm_peer_close": (ci, [ci, vp]),
        "vrm_peer_free": (ci, [ci, vp]),
        "vrm_copy_device": (ci, [ci, vp, vp, u64]),
        "vrm_set_l2_persistence": (ci, [vp, ci]),
        "vrm_set_statistics": (ci, [vp, ci]),
        "vrm_get_statistics": (ci, [vp, vp]),
        "vrm_microbench_l2": (ci, [ci, u64, C.POINTER(f32), C.POINTER(f32), C.POINTER(f32)]),
        "vrm_scene_set_completion_flag": (ci, [vp, vp, u32]),
        "vrm_scene_set_completion_counter": (ci, [vp, vp]),
        "vrm_claim_next": (ci, [ci, vp, vp, vp]),
        "vrm_wait_flags_device": (ci, [ci, vp, vp, u32, u32, u32, u32, vp]),
        "vrm_render_views_sharded": (ci, [vp, u32, vp, u32, vp, u32, ci, u32, u32, vp, ci, vp, C.POINTER(f32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def device_available() -> bool:
    return bool(load_library().vrm_device_available())


def microbench_l2(device: int = 0, working_set_bytes: int = 32 << 20):
    """L2 roofs for an L2-resident working set: dict(gather8_loads_per_ns, gather8_gbs, stream_gbs).  Measurement only."""
    a, b, c = C.c_float(), C.c_float(), C.c_float()
    rc = load_library().vrm_microbench_l2(device, working_set_bytes, C.byref(a), C.byref(b), C.byref(c))
    if rc:
        raise VrmError(f"vrm_microbench_l2 failed: {rc}")
    return dict(working_set_bytes=working_set_bytes, gather8_loads_per_ns=a.value, gather8_gbs=b.value, stream_gbs=c.value)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, np.float32).reshape(3))


def make_unit_vector(v):
    out = np.zeros(3, np.float32)
    rc = load_library().vrm_make_unit_vector(_ptr(_f3(v)), _ptr(out))
    if rc:
        raise VrmError("vrm_make_unit_vector failed")
    return out


class Camera:
    """The reference's pinhole camera (Camera.cuh:11-23).  ``data`` is the 60-byte struct as 15 float32:
    origin, lowerLeftCorner, horizontalVector, verticalVector, forwardVector."""

    def __init__(self, origin, look_at, up=(0.0, 1.0, 0.0), field_of_view=60.0, aspect_ratio=1920.0 / 1080.0):
        self.data = np.zeros(15, np.float32)
        rc = load_library().vrm_camera_make(_ptr(_f3(origin)), _ptr(_f3(look_at)), _ptr(_f3(up)),
                                            C.c_float(np.float32(field_of_view)), C.c_float(np.float32(aspect_ratio)), _ptr(self.data))
        if rc:
            raise VrmError("vrm_camera_make failed")

    @staticmethod
    def reference_default(width=1920, height=1080):
        """main/Main.cu:195-199."""
        return Camera((6.0, 2.0, 6.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 60.0, np.float32(width) / np.float32(height))


class VoxelScene:
    """One scene on one GPU (one handle = one device + one stream)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.vrm_scene_create(device, C.byref(self.h))
        if rc:
            self.h = C.c_void_p()
            raise VrmError(f"vrm_scene_create(device={device}) failed: {self.lib.vrm_error_string(rc).decode()} "
                           "(a CUDA device is required; there is no CPU fallback)")
        self.device = device
        self.storage = None
        self.build_ms = None

    # -- plumbing ------------------------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc:
            raise VrmError(f"{what}: {self.lib.vrm_error_string(rc).decode()}: {self.lib.vrm_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.vrm_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int):
        """Run on the caller's cudaStream_t (0 = the legacy default stream)."""
        self._check(self.lib.vrm_scene_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "vrm_scene_set_stream")

    def reset_stream(self):
        self._check(self.lib.vrm_scene_reset_stream(self.h), "vrm_scene_reset_stream")

    def synchronize(self):
        self._check(self.lib.vrm_scene_synchronize(self.h), "vrm_scene_synchronize")

    # -- scene construction --------------------------------------------------------------------------------------
    def insert_voxel(self, x: int, y: int, z: int, color: int):
        """``VoxelSceneCPU::insertVoxel``: one voxel per call (collected on the host by the library, no device work per call)."""
        rc = self.lib.vrm_scene_insert_voxel(self.h, x, y, z, color)
        if rc:
            self._check(rc, "vrm_scene_insert_voxel")

    def add_voxels(self, xyz, rgb):
        xyz = np.ascontiguousarray(xyz, np.int32).reshape(-1, 3)
        rgb = np.ascontiguousarray(rgb, np.uint32).reshape(-1)
        if xyz.shape[0] != rgb.shape[0]:
            raise ValueError("xyz and rgb disagree on the voxel count")
        self._check(self.lib.vrm_scene_add_voxels(self.h, _ptr(xyz), _ptr(rgb), xyz.shape[0]), "vrm_scene_add_voxels")

    def add_voxels_device(self, d_xyz_ptr: int, d_rgb_ptr: int, n: int):
        self._check(self.lib.vrm_scene_add_voxels_device(self.h, C.c_void_p(d_xyz_ptr), C.c_void_p(d_rgb_ptr), n), "vrm_scene_add_voxels_device")

    def generate_terrain(self, size: int = 512, seed: int = 1234, max_height: int = 0) -> int:
        """``scenes.terrain(size, seed, max_height)`` generated on the GPU straight into the staging list; returns the voxel count."""
        n = C.c_uint64()
        self._check(self.lib.vrm_scene_generate_terrain(self.h, size, seed, max_height, C.byref(n)), "vrm_scene_generate_terrain")
        return int(n.value)

    def generate_sparse_shells(self, size: int = 1024, cell: int = 64, seed: int = 7, fill_pct: int = 35) -> int:
        """``scenes.sparse_shells(size, cell, seed, fill_pct)`` generated on the GPU straight into the staging list."""
        n = C.c_uint64()
        self._check(self.lib.vrm_scene_generate_sparse_shells(self.h, size, cell, seed, fill_pct, C.byref(n)), "vrm_scene_generate_sparse_shells")
        return int(n.value)

    def generate_cube(self, x: int, y: int, z: int, half_width: int) -> int:
        """``VoxelCube::generateVoxelCube`` (= ``scenes.hollow_cube``) on the GPU, in the reference's insertion order."""
        n = C.c_uint64()
        self._check(self.lib.vrm_scene_generate_cube(self.h, x, y, z, half_width, C.byref(n)), "vrm_scene_generate_cube")
        return int(n.value)

    def generate_sphere(self, x: int, y: int, z: int, radius: int, checkered: bool = False) -> int:
        """``VoxelSphere::generate[Checkered]VoxelSphere`` (= ``scenes.sphere_shell``) on the GPU, in the reference's insertion order."""
        n = C.c_uint64()
        self._check(self.lib.vrm_scene_generate_sphere(self.h, x, y, z, radius, int(checkered), C.byref(n)), "vrm_scene_generate_sphere")
        return int(n.value)

    def generate_voxel_scene(self, storage_type):
        st = STORAGE[storage_type] if isinstance(storage_type, str) else int(storage_type)
        ms = C.c_float()
        self._check(self.lib.vrm_scene_build(self.h, st, C.byref(ms)), "vrm_scene_build")
        self.storage = st
        self.build_ms = ms.value
        return ms.value

    def info(self):
        d, m, f, u, b = C.c_uint32(), C.c_int32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.vrm_scene_info(self.h, C.byref(d), C.byref(m), C.byref(f), C.byref(u), C.byref(b)), "vrm_scene_info")
        return dict(diameter=d.value, min_coord=m.value, filled=f.value, unique_voxels=u.value, bytes=b.value)

    def get_array_diameter(self):
        return self.info()["diameter"]

    def get_array_size(self):
        return self.info()["diameter"] ** 3

    def get_min_coord(self):
        return self.info()["min_coord"]

    # -- lighting ------------------------------------------------------------------------------------------------
    def setup_constant_values(self, light_direction=None, light_color=(1.0, 1.0, 1.0), light_position=(10.0, 10.0, -10.0),
                              use_point_light=False, use_shadows=True):
        if light_direction is None:
            light_direction = make_unit_vector((1.0, 1.0, 1.0))
        self._check(self.lib.vrm_set_lighting(self.h, _ptr(_f3(light_direction)), _ptr(_f3(light_color)), _ptr(_f3(light_position)),
                                              int(use_point_light), int(use_shadows)), "vrm_set_lighting")

    # -- rendering -----------------------------------------------------------------------------------------------
    @staticmethod
    def _algo(algorithm):
        return ALGORITHM[algorithm] if isinstance(algorithm, str) else int(algorithm)

    @staticmethod
    def _cam(camera):
        return np.ascontiguousarray(camera.data if isinstance(camera, Camera) else camera, np.float32)

    def run_raymarching_kernel(self, width, height, algorithm, camera, scale=1, translation=(0.0, 0.0, 0.0), want_hits=False,
                               rgb_out=None):
        """Render one frame into host memory.  Returns dict(rgb[H,W,3] uint8, hits[H,W,4] int32 | None, kernel_ms)."""
        rgb = rgb_out if rgb_out is not None else np.zeros((height, width, 3), np.uint8)
        hits = np.zeros((height, width, 4), np.int32) if want_hits else None
        ms = C.c_float()
        cam = self._cam(camera)
        self._check(self.lib.vrm_render(self.h, _ptr(cam), _ptr(_f3(translation)), scale, self._algo(algorithm), width, height,
                                        _ptr(rgb), _ptr(hits), C.byref(ms)), "vrm_render")
        return dict(rgb=rgb, hits=hits, kernel_ms=ms.value)

    render = run_raymarching_kernel

    def render_device(self, width, height, algorithm, camera, d_rgb_ptr: int, d_hits_ptr: int | None = None, scale=1,
                      translation=(0.0, 0.0, 0.0)):
        cam = self._cam(camera)
        self._check(self.lib.vrm_render_device(self.h, _ptr(cam), _ptr(_f3(translation)), scale, self._algo(algorithm), width, height,
                                               C.c_void_p(d_rgb_ptr), C.c_void_p(d_hits_ptr or 0)), "vrm_render_device")

    def render_views_device(self, width, height, algorithm, cameras, d_rgb_ptr: int, d_hits_ptr: int | None = None, scale=1,
                            translation=(0.0, 0.0, 0.0), view_stride: int = 1):
        """All ``cameras`` in one launch; view ``v`` goes to frame slot ``v * view_stride`` of the outputs."""
        cams = np.ascontiguousarray(np.stack([self._cam(c) for c in cameras]), np.float32)
        self._check(self.lib.vrm_render_views_device_strided(self.h, _ptr(cams), cams.shape[0], _ptr(_f3(translation)), scale, self._algo(algorithm),
                                                             width, height, C.c_void_p(d_rgb_ptr), C.c_void_p(d_hits_ptr or 0), int(view_stride)),
                    "vrm_render_views_device_strided")

    def render_views(self, width, height, algorithm, cameras, scale=1, translation=(0.0, 0.0, 0.0), rgb_out=None):
        """A batch of views (camera orbit) into host frames ``[n, H, W, 3]``; ``rgb_out`` may be a pinned buffer (written directly
        by the kernel) or pageable memory (double-buffered device batches, copies overlapped with rendering)."""
        cams = np.ascontiguousarray(np.stack([self._cam(c) for c in cameras]).astype(np.float32))
        n = cams.shape[0]
        if rgb_out is None:
            rgb_out = np.zeros((n, height, width, 3), np.uint8)
        assert rgb_out.dtype == np.uint8 and rgb_out.size == n * height * width * 3 and rgb_out.flags["C_CONTIGUOUS"]
        ms = C.c_float()
        self._check(self.lib.vrm_render_views(self.h, _ptr(cams), n, _ptr(_f3(translation)), int(scale), self._algo(algorithm), width, height,
                                              _ptr(rgb_out), C.byref(ms)), "vrm_render_views")
        return {"rgb": rgb_out.reshape(n, height, width, 3), "total_ms": ms.value}

    def trace_rays(self, rays, algorithm, scale=1, translation=(0.0, 0.0, 0.0), want_hits=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        colour = np.zeros(n, np.uint32)
        hits = np.zeros((n, 4), np.int32) if want_hits else None
        ms = C.c_float()
        self._check(self.lib.vrm_trace_rays(self.h, _ptr(rays), n, _ptr(_f3(translation)), scale, self._algo(algorithm), _ptr(colour),
                                            _ptr(hits), C.byref(ms)), "vrm_trace_rays")
        return dict(colour=colour, hits=hits, kernel_ms=ms.value)

    def trace_rays_device(self, d_rays_ptr: int, n: int, algorithm, d_colour_ptr: int, d_hits_ptr: int | None = None, scale=1,
                          translation=(0.0, 0.0, 0.0)):
        self._check(self.lib.vrm_trace_rays_device(self.h, C.c_void_p(d_rays_ptr), n, _ptr(_f3(translation)), scale, self._algo(algorithm),
                                                   C.c_void_p(d_colour_ptr), C.c_void_p(d_hits_ptr or 0)), "vrm_trace_rays_device")

    def lookup(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.int32).reshape(-1, 3)
        n = xyz.shape[0]
        out = np.zeros(n, np.uint32)
        exists = np.zeros(n, np.uint8)
        self._check(self.lib.vrm_lookup(self.h, _ptr(xyz), n, _ptr(out), _ptr(exists)), "vrm_lookup")
        return out, exists

    def set_completion_flag(self, d_flag_ptr: int | None, first_value: int = 1):
        """Every later render launch of this scene stores first_value, first_value + 1, ... into the word at ``d_flag_ptr`` (device or
        peer memory) behind its frame (``multigpu.PeerFrameBuffer.flag_ptr``); ``None`` switches the signal off."""
        self._check(self.lib.vrm_scene_set_completion_flag(self.h, C.c_void_p(d_flag_ptr or 0), first_value), "vrm_scene_set_completion_flag")

    def set_completion_counter(self, d_counter_ptr: int | None):
        """Counting form of the completion signal: every render launch adds its number of views to the word at ``d_counter_ptr``."""
        self._check(self.lib.vrm_scene_set_completion_counter(self.h, C.c_void_p(d_counter_ptr or 0)), "vrm_scene_set_completion_counter")

    # -- statistics ----------------------------------------------------------------------------------------------
    def set_l2_persistence(self, enabled: bool):
        self._check(self.lib.vrm_set_l2_persistence(self.h, int(enabled)), "vrm_set_l2_persistence")

    def set_statistics(self, enabled: bool, as_executed: bool = False):
        """Event counters for the next render / trace calls.  Default: comparable with the reference's (every shadow ray is traced, as
        the reference does).  ``as_executed``: the counters of the work the production kernels do -- a shadow ray whose pixel is already
        black (normal facing away from the light: colour * !shadow = 0 either way) is not traced."""
        self._check(self.lib.vrm_set_statistics(self.h, (2 if as_executed else 1) if enabled else 0), "vrm_set_statistics")

    def get_statistics(self):
        out = np.zeros(8, np.uint64)
        self._check(self.lib.vrm_get_statistics(self.h, _ptr(out)), "vrm_get_statistics")
        keys = ["exist_checks", "exist_false", "lookups", "lookup_hits", "table2_probes", "region_reads", "rays", "crawl_skipped"]
        return {k: int(v) for k, v in zip(keys, out)}


def render_views_sharded(scene_list, width, height, algorithm, cameras, scale=1, translation=(0.0, 0.0, 0.0), rgb_out=None, d_rgb_ptr: int | None = None):
    """``vrm_render_views_sharded``: ONE process, one built ``VoxelScene`` per device (same voxels); the views are claimed dynamically
    and gathered on the first scene's device.  Returns dict(rgb[n,H,W,3] | None, views_per_scene, total_ms); with ``d_rgb_ptr`` the
    frames stay in that device buffer (on the first scene's device)."""
    lib = load_library()
    cams = np.ascontiguousarray(np.stack([VoxelScene._cam(c) for c in cameras]).astype(np.float32))
    n = cams.shape[0]
    handles = (C.c_void_p * len(scene_list))(*[s.h for s in scene_list])
    counts = np.zeros(len(scene_list), np.uint32)
    ms = C.c_float()
    on_device = d_rgb_ptr is not None
    if not on_device:
        if rgb_out is None:
            rgb_out = np.zeros((n, height, width, 3), np.uint8)
        assert rgb_out.dtype == np.uint8 and rgb_out.size == n * height * width * 3 and rgb_out.flags["C_CONTIGUOUS"]
    out = C.c_void_p(d_rgb_ptr) if on_device else _ptr(rgb_out)
    rc = lib.vrm_render_views_sharded(C.cast(handles, C.c_void_p), len(scene_list), _ptr(cams), n, _ptr(_f3(translation)), int(scale),
                                      VoxelScene._algo(algorithm), width, height, out, int(on_device), _ptr(counts), C.byref(ms))
    if rc:
        raise VrmError(f"vrm_render_views_sharded: {lib.vrm_error_string(rc).decode()}: {lib.vrm_last_error(scene_list[0].h).decode()}")
    return dict(rgb=None if on_device else rgb_out.reshape(n, height, width, 3), views_per_scene=counts.tolist(), total_ms=ms.value)
