"""Where does the end-to-end time of vrm_render go (one GPU, the bench workload)?  For each form of the page-locked path the wall time of
the call (what bench.py's `e2e` measures), the event time the call reports (camera upload excluded, kernels + frame landing included)
and the device-frame kernel time beside them.

    python tools/e2e_breakdown.py [--iters 20] [--forms default,dma2,dma4,wstore]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxelraymarcher_b200 import api  # noqa: E402
import bench  # noqa: E402

FORMS = {
    "default": {},
    "dma1": {"VRM_PINNED_DMA": "1"},
    "dma2": {"VRM_PINNED_DMA": "2"},
    "dma4": {"VRM_PINNED_DMA": "4"},
    "dma8": {"VRM_PINNED_DMA": "8"},
    "wstore": {"VRM_WSTORE_REMOTE": "1"},
    "nobulk": {"VRM_BULK_STORE": "0"},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--forms", default="default,dma1,dma2,dma4,wstore")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    W, H = bench.WIDTH, bench.HEIGHT
    flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
    host = torch.zeros((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
    frame = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    cams = [bench.orbit_camera(api, bench.view_of(i, 0, 1)) for i in range(a.iters + 3)]
    for name in a.forms.split(","):
        for k in ("VRM_PINNED_DMA", "VRM_WSTORE_REMOTE", "VRM_BULK_STORE"):
            os.environ.pop(k, None)
        os.environ.update(FORMS[name])
        s = api.VoxelScene(0)
        s.set_stream(stream.cuda_stream)
        s.generate_terrain(bench.SCENE_SIZE, bench.SCENE_SEED)
        s.generate_voxel_scene("vcs")
        wall, ev, dk = [], [], []
        same = True
        for i, cam in enumerate(cams):
            flush.add_(1)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            r = s.render(W, H, bench.ALGORITHM, cam, rgb_out=host)
            dt = time.perf_counter() - t0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.add_(1)
            e0.record(stream)
            s.render_device(W, H, bench.ALGORITHM, cam, frame.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize(dev)
            same = same and bool(np.array_equal(host, frame.cpu().numpy()))
            host[:] = 0
            if i >= 3:
                wall.append(dt * 1e3); ev.append(r["kernel_ms"]); dk.append(e0.elapsed_time(e1))
        rays = W * H / 1e6
        print(f"{name:8s} wall {np.mean(wall):.3f} ms ({rays / np.mean(wall) * 1e3:.0f} Mrays/s)  event time in the call {np.mean(ev):.3f}  "
              f"host overhead {np.mean(wall) - np.mean(ev):.3f}  device-frame kernel {np.mean(dk):.3f}  frames equal {same}", flush=True)
        s.close()


if __name__ == "__main__":
    main()
