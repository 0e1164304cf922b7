"""Build-time probe on a GPU box: 512^3 terrain generated on the GPU, built as VCS and as hash table, three times each
(VRM_BUILD_TRACE=1 in the environment prints the builder's stage marks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelraymarcher_b200 import api
for storage in ("vcs", "hashtable"):
    for i in range(3):
        s = api.VoxelScene(0)
        t0 = time.perf_counter(); n = s.generate_terrain(512, 1234); t1 = time.perf_counter()
        ms = s.generate_voxel_scene(storage); t2 = time.perf_counter()
        print(f"{storage} run {i}: generate {1e3 * (t1 - t0):.1f} ms ({n} voxels), build event {ms:.2f} ms, build wall {1e3 * (t2 - t1):.1f} ms, {n / ms / 1e3:.0f} Mvoxels/s", flush=True)
        s.close()
