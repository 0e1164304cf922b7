"""Config-5 timing only (no reference arms): 1024^3 sparse shells generated on the GPU, 8.3 M incoherent rays + shadows,
VCS, both algorithms; median kernel time of 7 launches.  For A/B builds (VRM_B200_LIB)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelraymarcher_b200 import api, scenes
n = 3840 * 2160
rays = scenes.random_rays(n, (512.0 + 31.5, 512.0 + 31.5, 512.0 + 31.5), seed=42)
s = api.VoxelScene(0)
s.set_stream(torch.cuda.current_stream().cuda_stream)
s.generate_sparse_shells(1024, 64, 11, 35)
s.generate_voxel_scene("vcs")
d_rays = torch.from_numpy(rays).cuda()
d_col = torch.zeros(n, dtype=torch.int32, device="cuda:0")
for algo in ("longestaxis", "original"):
    ts = []
    for i in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.trace_rays_device(d_rays.data_ptr(), n, algo, d_col.data_ptr()); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    print(f"{os.environ.get('VRM_B200_LIB', 'main').split('libvrm_')[-1]:14s} trace {algo:12s} {np.median(ts):8.3f} ms  hit fraction {float((d_col != 0).float().mean()):.5f}", flush=True)
