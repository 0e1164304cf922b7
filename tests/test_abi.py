"""CPU tier: the C-ABI library loads, exports every symbol include/vrm_b200.h declares, its host-side helpers agree with
the oracle, and compute entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.common import PROBE_CAMERAS, ROOT, camera, po
from voxelraymarcher_b200 import api


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vrm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vrm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(api.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    assert sorted(names) == sorted(api.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_camera_and_unit_vector_match_oracle():
    for o, l, fov in PROBE_CAMERAS:
        for w, h in ((1280, 720), (1920, 1080), (3840, 2160), (160, 90)):
            mine = api.Camera(o, l, (0.0, 1.0, 0.0), fov, np.float32(w) / np.float32(h)).data
            assert np.array_equal(mine, camera(o, l, fov, w, h, "orc"))
            if po.available("refh"):
                assert np.array_equal(mine, camera(o, l, fov, w, h, "refh"))
    assert np.array_equal(api.make_unit_vector((1, 1, 1)), po.unit_vector((1, 1, 1), "orc"))
    assert np.array_equal(api.Camera.reference_default().data, camera((6, 2, 6), (0, 0, -1), 60.0, 1920, 1080, "orc"))


def test_error_strings_and_invalid_arguments():
    lib = api.load_library()
    assert lib.vrm_error_string(0) == b"ok"
    assert lib.vrm_error_string(2) == b"CUDA error"
    assert lib.vrm_scene_create(0, None) == 1          # VRM_ERR_INVALID
    assert lib.vrm_scene_destroy(None) == 0
    assert lib.vrm_last_error(None) == b""


@pytest.mark.skipif(api.device_available(), reason="this check is for the CPU-only tier")
def test_no_cpu_fallback_without_a_device():
    lib = api.load_library()
    h = C.c_void_p()
    assert lib.vrm_scene_create(0, C.byref(h)) == 2     # VRM_ERR_CUDA: fails loudly
    assert not h.value
    with pytest.raises(api.VrmError):
        api.VoxelScene(0)
