"""Exploratory timing on a GPU box (not the bench): ours vs the reference's own CUDA kernels on a chosen scene."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelraymarcher_b200 import api, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="terrain")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--ref", action="store_true", help="also time the reference's CUDA kernels (slow host-side build)")
    ap.add_argument("--combos", default="hashtable:original,hashtable:longestaxis,vcs:original,vcs:longestaxis")
    ap.add_argument("--out", default="gpurun_out/explore.json")
    a = ap.parse_args()
    t0 = time.time()
    if a.scene == "terrain":
        xyz, rgb = scenes.terrain(a.size, 1234)
        scale = 1
        org, look = (-96.0, 352.0, -96.0), (256.0, 64.0, 256.0)
        if a.size != 512:
            f = a.size / 512.0
            org, look = tuple(v * f for v in org), tuple(v * f for v in look)
    else:
        xyz, rgb = scenes.probe_scene()
        scale = 8
        org, look = (6.0, 2.0, 6.0), (0.0, 0.0, -1.0)
    print(f"scene {a.scene}: {xyz.shape[0]} voxels generated in {time.time() - t0:.1f}s", flush=True)
    W, H = a.width, a.height
    cam = api.Camera(org, look, (0.0, 1.0, 0.0), 60.0, np.float32(W) / np.float32(H))
    results = []
    combos = [c.split(":") for c in a.combos.split(",")]
    for storage in sorted({c[0] for c in combos}):
        s = api.VoxelScene(0)
        t0 = time.time()
        s.add_voxels(xyz, rgb)
        t1 = time.time()
        ms = s.generate_voxel_scene(storage)
        info = s.info()
        print(f"[{storage}] h2d {t1 - t0:.2f}s build {ms:.1f} ms ({xyz.shape[0] / ms / 1e3:.1f} Mvoxels/s) info {info}", flush=True)
        ref = None
        if a.ref:
            from oracle import pyoracle as po
            po.set_lighting("refg")
            t0 = time.time()
            ref = po.OracleScene("refg")
            ref.add_voxels(xyz, rgb)
            ref.build(storage)
            print(f"[{storage}] reference host-side build {time.time() - t0:.1f}s", flush=True)
        for st, algo in combos:
            if st != storage:
                continue
            # device-resident frame, CUDA events on the launching stream, L2 flushed between launches (as bench.py does); the first
            # render goes through the host path once for the hit fraction
            import torch
            r = s.render(W, H, algo, cam, scale=scale, want_hits=True)
            hitfrac = float(r["hits"][..., 3].mean())
            cur = torch.cuda.current_stream()
            s.set_stream(cur.cuda_stream)
            fb = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda:0")
            flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda:0")
            times = []
            for i in range(a.iters + 2):
                flush.add_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur); s.render_device(W, H, algo, cam, fb.data_ptr(), scale=scale); e1.record(cur)
                torch.cuda.synchronize()
                if i >= 2:
                    times.append(e0.elapsed_time(e1))
            del flush
            s.reset_stream()
            s.set_statistics(True)
            s.render(W, H, algo, cam, scale=scale)
            stats = s.get_statistics()
            s.set_statistics(False)
            med = float(np.median(times))
            row = dict(storage=storage, algo=algo, ms=med, mrays=W * H / med / 1e3, hit_fraction=hitfrac, stats=stats, build_ms=ms, bytes=info["bytes"])
            if ref is not None:
                import ctypes as C
                msout = np.zeros(a.iters, np.float32)
                rc = ref.lib.refg_render_timed(ref.h, po._ptr(cam.data), po._ptr(np.zeros(3, np.float32)), scale, po.ALGORITHM[algo], W, H, 2, a.iters, po._ptr(msout))
                rmed = float(np.median(msout))
                row.update(ref_ms=rmed, ref_mrays=W * H / rmed / 1e3, speedup=rmed / med, ref_rc=rc)
                rr = ref.render(cam.data, W, H, algo, scale=scale, want_hits=False)
                close = float((np.abs(r["rgb"].astype(np.int32) - rr["rgb"].astype(np.int32)) <= 1).all(-1).mean())
                row.update(rgb_within_1lsb_vs_ref_fmad=close)
            print(json.dumps(row), flush=True)
            results.append(row)
        s.close()
        if ref is not None:
            ref.close()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(results, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
