"""The reference's process-level seam: `VoxelRaymarcher <scale> <hashtable|vcs> <original|longestaxis>` (Main.cu:176-229)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests.common import ROOT, scenes
from voxelraymarcher_b200 import api

CLI = os.path.join(ROOT, "voxelraymarcher_b200", "VoxelRaymarcher")


def test_cli_requires_a_scale():
    p = subprocess.run([CLI], stdout=subprocess.PIPE, text=True)
    assert p.returncode == 1                                  # Main.cu:181-185
    assert p.stdout.strip() == "You need to provide a voxel scale"


@pytest.mark.skipif(api.device_available(), reason="CPU-only tier check")
def test_cli_fails_loudly_without_a_gpu(tmp_path):
    p = subprocess.run([CLI, "8", "hashtable", "original"], stdout=subprocess.PIPE, text=True, cwd=tmp_path)
    assert p.returncode != 0
    assert "Storage Type: Cuckoo Hash Table" in p.stdout and "Raymarching Algorithm: Original" in p.stdout
    assert "Device Count: 0" in p.stdout and "ERROR" in p.stdout


def _run_cli(tmp_path, args):
    p = subprocess.run([CLI, *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=tmp_path)
    assert p.returncode == 0, p.stdout
    from PIL import Image
    return p.stdout, np.asarray(Image.open(os.path.join(tmp_path, "output.png")).convert("RGB"))


@pytest.mark.gpu
@pytest.mark.parametrize("argv,storage,algo", [(["8", "hashtable", "original"], "hashtable", "original"), (["8"], "vcs", "longestaxis"),
                                               (["8", "vcs", "bogus"], "vcs", "longestaxis")])
def test_cli_matches_api_on_csv_scene(tmp_path, argv, storage, algo):
    xyz, rgb = scenes.probe_scene()
    os.makedirs(tmp_path / "resources")
    scenes.write_csv(str(tmp_path / "resources" / "scene.vox"), xyz, rgb)
    out, img = _run_cli(tmp_path, argv)
    for line in ("Device Count:", "regions that are filled", "Storage Structures Generated", "Execution Time for Ray Marching Algorithm is:"):
        assert line in out
    assert "There are : 13/125 regions that are filled" in out
    s = api.VoxelScene(0)
    s.add_voxels(xyz, rgb)
    s.generate_voxel_scene(storage)
    want = s.render(1920, 1080, algo, api.Camera.reference_default(), scale=8)["rgb"]
    assert img.shape == (1080, 1920, 3)
    assert np.array_equal(img, want)


@pytest.mark.gpu
def test_cli_reads_magicavoxel(tmp_path):
    rng = np.random.default_rng(3)
    pts = np.unique(rng.integers(0, 40, size=(3000, 3)), axis=0).astype(np.uint8)
    idx = rng.integers(1, 256, size=pts.shape[0]).astype(np.uint8)
    palette = rng.integers(0, 256, size=(256, 4)).astype(np.uint8)

    def chunk(cid, body, children=b""):
        return cid + struct.pack("<II", len(body), len(children)) + body + children
    xyzi = struct.pack("<I", pts.shape[0]) + np.concatenate([pts, idx[:, None]], 1).tobytes()
    kids = chunk(b"SIZE", struct.pack("<III", 40, 40, 40)) + chunk(b"XYZI", xyzi) + chunk(b"RGBA", palette.tobytes())
    os.makedirs(tmp_path / "resources")
    with open(tmp_path / "resources" / "scene.vox", "wb") as f:
        f.write(b"VOX " + struct.pack("<I", 150) + chunk(b"MAIN", b"", kids))
    _, img = _run_cli(tmp_path, ["1", "vcs", "original", "--width", "640", "--height", "360"])
    xyz = np.stack([pts[:, 0], pts[:, 2], pts[:, 1]], 1).astype(np.int32)      # z-up -> y-up
    pal = palette[idx.astype(np.int32) - 1].astype(np.uint32)                  # palette entry i is stored at i-1
    rgb = (pal[:, 0] << 16) | (pal[:, 1] << 8) | pal[:, 2]
    s = api.VoxelScene(0)
    s.add_voxels(xyz, rgb)
    s.generate_voxel_scene("vcs")
    cam = api.Camera((6.0, 2.0, 6.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 60.0, np.float32(640) / np.float32(360))
    want = s.render(640, 360, "original", cam, scale=1, want_hits=True)
    assert np.array_equal(img, want["rgb"])
    assert want["hits"][..., 3].sum() > 1000
