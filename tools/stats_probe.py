"""Event counters + kernel time of one view of config 4 (exploration)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelraymarcher_b200 import api, scenes
xyz, rgb = scenes.sparse_shells(2048, 64, seed=7, fill_pct=35)
w, h = 1920, 1080
s = api.VoxelScene(0); s.add_voxels(xyz, rgb); s.generate_voxel_scene("vcs")
for v in (0, 5, 11):
    ang = 2.0 * np.pi * (v + 0.37) / 64
    r, el = 1.5 * 1024.0, np.deg2rad(20.0)
    org = (float(1024 + r * np.cos(el) * np.cos(ang)), float(1024 + r * np.sin(el)), float(1024 + r * np.cos(el) * np.sin(ang)))
    cam = api.Camera(org, (1024.0, 1024.0, 1024.0), (0.0, 1.0, 0.0), 60.0, np.float32(w) / np.float32(h))
    for algo in ("original", "longestaxis"):
        s.set_statistics(False)
        ms = [s.render(w, h, algo, cam)["kernel_ms"] for _ in range(3)][-1]
        s.set_statistics(True)
        s.render(w, h, algo, cam)
        st = s.get_statistics()
        print(v, algo, "ms", round(ms, 3), {k: round(val / st["rays"], 2) for k, val in st.items() if k != "rays"}, flush=True)
