"""Wall-clock of the host-buffer call vrm_render on the bench workload for a given VRM_BANDS (exploration)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from voxelraymarcher_b200 import api, scenes
xyz, rgb = scenes.terrain(512, 1234)
s = api.VoxelScene(0); s.add_voxels(xyz, rgb); s.generate_voxel_scene("vcs")
cam = bench.orbit_camera(api, 0)
host = torch.zeros((2160, 3840, 3), dtype=torch.uint8).pin_memory().numpy()
flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda:0")
ts, ks = [], []
for i in range(12):
    flush.add_(1); torch.cuda.synchronize()
    t0 = time.perf_counter(); r = s.render(3840, 2160, "longestaxis", cam, rgb_out=host); dt = time.perf_counter() - t0
    if i >= 3: ts.append(dt * 1e3); ks.append(r["kernel_ms"])
print("e2e ms median", round(float(np.median(ts)), 3), "kernel-span ms", round(float(np.median(ks)), 3))
