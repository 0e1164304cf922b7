"""Small single-purpose workloads for ncu captures of the kernels bench.py does not launch:

    python tools/ncu_targets.py trace5      config 5: 8.29 M incoherent rays on the 1024^3 sparse scene, VCS + longest axis (trace_kernel)
    python tools/ncu_targets.py orbit4      config 4: one 1080p view of the 2048^3 orbit, VCS + longest axis (render_kernel on a 900 MB structure)
    python tools/ncu_targets.py build3 [vcs|hashtable]   config 3: the 32 M voxel build (GPU-generated terrain, so the process is short)

Each runs the launches a few times (ncu: -k regex:<kernel> -s <skip> -c 1) and prints the event time of the last run."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from voxelraymarcher_b200 import api, scenes  # noqa: E402


def timed(fn, reps=3):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms


def main():
    target = sys.argv[1]
    s = api.VoxelScene(0)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    if target == "trace5":
        s.generate_sparse_shells(1024, 64, 11, 35)
        s.generate_voxel_scene("vcs")
        n = 3840 * 2160
        rays = torch.from_numpy(scenes.random_rays(n, (512.0 + 31.5, 512.0 + 31.5, 512.0 + 31.5), seed=42)).cuda()
        col = torch.zeros(n, dtype=torch.int32, device="cuda:0")
        algo = sys.argv[2] if len(sys.argv) > 2 else "longestaxis"
        print("trace5", algo, timed(lambda: s.trace_rays_device(rays.data_ptr(), n, algo, col.data_ptr())))
    elif target == "orbit4":
        t0 = time.time()
        s.generate_sparse_shells(2048, 64, 7, 35)
        ms = s.generate_voxel_scene("vcs")
        print(f"2048^3 scene built in {ms:.1f} ms (+ generation), {time.time() - t0:.1f}s wall", s.info())
        w, h = 1920, 1080
        from tests.test_configs_gpu import orbit_camera_2048
        fb = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda:0")
        algo = sys.argv[2] if len(sys.argv) > 2 else "longestaxis"
        for v in (19, 38):
            cam = orbit_camera_2048(v, w, h)
            print("orbit4 view", v, algo, timed(lambda: s.render_device(w, h, algo, cam, fb.data_ptr())))
    elif target == "build3":
        storage = sys.argv[2] if len(sys.argv) > 2 else "vcs"
        for rep in range(2):          # the second build finds its scratch in the memory pool
            t = api.VoxelScene(0)
            n = t.generate_terrain(512, 1234)
            t.synchronize()
            t0 = time.perf_counter()
            ms = t.generate_voxel_scene(storage)
            wall = (time.perf_counter() - t0) * 1e3
            print(f"build3 {storage} rep {rep}: {n} voxels, event {ms:.2f} ms, host wall {wall:.2f} ms, {n / ms / 1e3:.0f} Mvoxels/s", t.info())
            t.close()
    else:
        raise SystemExit(f"unknown target {target}")
    s.close()


if __name__ == "__main__":
    main()
