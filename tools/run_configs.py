"""Runs the five BASELINE.json configurations on a GPU box and writes one JSON report (gpurun_out/configs.json):
ours (kernel time, median of N launches, L2 flushed), the reference's own CUDA kernels rebuilt for sm_100a (same scene,
same camera / rays, default nvcc flags) and the reference's host build on all cores, plus parity figures where the CPU
reference finishes quickly.  This is the reporting grid of BASELINE.md §4; bench.py remains the contract benchmark."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import pyoracle as po  # noqa: E402
from voxelraymarcher_b200 import api, scenes  # noqa: E402

ZERO3 = np.zeros(3, np.float32)


def time_ours(scene, fn, iters, flush):
    ts = []
    for i in range(iters + 2):
        flush.add_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def ref_gpu_render_ms(ref, cam, scale, algo, w, h, iters):
    ms = np.zeros(iters, np.float32)
    rc = ref.lib.refg_render_timed(ref.h, po._ptr(cam), po._ptr(ZERO3), scale, po.ALGORITHM[algo], w, h, 2, iters, po._ptr(ms))
    return float(np.median(ms)) if rc == 0 else None


def frames_config(name, xyz, rgb, scale, cams, w, h, combos, iters, want_parity, flush, out):
    cores = os.cpu_count() or 1
    for storage in sorted({c[0] for c in combos}):
        s = api.VoxelScene(0)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        s.add_voxels(xyz, rgb)
        build_ms = s.generate_voxel_scene(storage)
        fb = torch.zeros((len(cams), h, w, 3), dtype=torch.uint8, device="cuda:0")
        refg = refh = None
        if po.available("refg"):
            po.set_lighting("refg")
            t0 = time.perf_counter()
            refg = po.OracleScene("refg"); refg.add_voxels(xyz, rgb); refg.build(storage)
            ref_build_s = time.perf_counter() - t0
        if want_parity and po.available("refh"):
            po.set_lighting("refh")
            refh = po.OracleScene("refh"); refh.add_voxels(xyz, rgb); refh.build(storage)
        for st, algo in combos:
            if st != storage:
                continue
            ms = time_ours(s, lambda: s.render_views_device(w, h, algo, cams, fb.data_ptr(), scale=scale), iters, flush)
            rays = w * h * len(cams)
            row = dict(config=name, storage=storage, algo=algo, views=len(cams), resolution=f"{w}x{h}", ms=ms, ms_per_frame=ms / len(cams),
                       mrays_per_s=rays / ms / 1e3, build_ms=build_ms, voxels=int(xyz.shape[0]), structure_bytes=s.info()["bytes"])
            if refg is not None:
                per = [ref_gpu_render_ms(refg, c.data, scale, algo, w, h, 3) for c in cams[: min(len(cams), 4)]]
                if all(p is not None for p in per):
                    rms = float(np.mean(per))
                    row.update(ref_gpu_ms_per_frame=rms, ref_gpu_mrays_per_s=w * h / rms / 1e3, speedup_vs_ref_gpu=rms / (ms / len(cams)), ref_host_build_s=ref_build_s)
            if refh is not None:
                got = s.render(w, h, algo, cams[0], scale=scale, want_hits=True)
                t0 = time.perf_counter()
                want = refh.render(cams[0].data, w, h, algo, scale=scale, threads=cores)
                cpu_s = time.perf_counter() - t0
                hm = float((got["hits"] == want["hits"]).all(-1).mean())
                close = float((np.abs(got["rgb"].astype(np.int32) - want["rgb"].astype(np.int32)) <= 1).all(-1).mean())
                row.update(hit_map_parity=hm, rgb_within_1lsb=close, ref_cpu_mrays_per_s=w * h / cpu_s / 1e6, ref_cpu_cores=cores,
                           hit_pixels=int(want["hits"][..., 3].sum()))
            print(json.dumps(row), flush=True)
            out.append(row)
        s.close()
        for r in (refg, refh):
            if r is not None:
                r.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/configs.json")
    a = ap.parse_args()
    which = {int(c) for c in a.configs.split(",")}
    flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda:0")
    out = []
    all4 = [(s, al) for s in ("hashtable", "vcs") for al in ("original", "longestaxis")]
    if which & {1, 2}:
        xyz, rgb = scenes.probe_scene()
        if 1 in which:
            cam = api.Camera((6.0, 2.0, 6.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 60.0, np.float32(1280) / np.float32(720))
            frames_config("1: scene.vox stand-in 1280x720", xyz, rgb, 8, [cam], 1280, 720, [("hashtable", "original")], a.iters, True, flush, out)
        if 2 in which:
            cam = api.Camera.reference_default(1920, 1080)
            frames_config("2: scene.vox stand-in 1920x1080", xyz, rgb, 8, [cam], 1920, 1080, all4, a.iters, True, flush, out)
    if 3 in which:
        xyz, rgb = scenes.terrain(512, 1234)
        cam = api.Camera((-96.0, 352.0, -96.0), (256.0, 64.0, 256.0), (0.0, 1.0, 0.0), 60.0, np.float32(3840) / np.float32(2160))
        frames_config("3: 512^3 terrain 3840x2160", xyz, rgb, 1, [cam], 3840, 2160, all4, a.iters, True, flush, out)
    if 4 in which:
        t0 = time.time()
        xyz, rgb = scenes.sparse_shells(2048, 64, seed=7, fill_pct=35)
        print(f"config 4 scene: {xyz.shape[0]} voxels in {time.time() - t0:.1f}s", flush=True)
        w, h = 1920, 1080
        cams = []
        for v in range(64):
            ang = 2.0 * np.pi * (v + 0.37) / 64           # offset: no view is exactly axis-aligned
            r, el = 1.5 * 1024.0, np.deg2rad(20.0)
            org = (float(1024 + r * np.cos(el) * np.cos(ang)), float(1024 + r * np.sin(el)), float(1024 + r * np.cos(el) * np.sin(ang)))
            cams.append(api.Camera(org, (1024.0, 1024.0, 1024.0), (0.0, 1.0, 0.0), 60.0, np.float32(w) / np.float32(h)))
        frames_config("4: 2048^3 sparse shells, 64-view orbit 1920x1080", xyz, rgb, 1, cams, w, h, [("vcs", "longestaxis"), ("vcs", "original")], 3, False, flush, out)
    if 5 in which:
        xyz, rgb = scenes.sparse_shells(1024, 64, seed=11, fill_pct=35)
        n = 3840 * 2160
        rays = scenes.random_rays(n, (512.0 + 31.5, 512.0 + 31.5, 512.0 + 31.5), seed=42)
        s = api.VoxelScene(0)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        s.add_voxels(xyz, rgb)
        build_ms = s.generate_voxel_scene("vcs")
        d_rays = torch.from_numpy(rays).cuda()
        d_col = torch.zeros(n, dtype=torch.int32, device="cuda:0")
        for algo in ("longestaxis", "original"):
            ms = time_ours(s, lambda: s.trace_rays_device(d_rays.data_ptr(), n, algo, d_col.data_ptr()), a.iters, flush)
            row = dict(config="5: 1024^3 sparse shells, incoherent rays + shadows, 4K ray count", storage="vcs", algo=algo, rays=n, ms=ms, mrays_per_s=n / ms / 1e3,
                       build_ms=build_ms, voxels=int(xyz.shape[0]), hit_fraction=float((d_col != 0).float().mean()))
            if po.available("refg"):
                po.set_lighting("refg")
                ref = po.OracleScene("refg"); ref.add_voxels(xyz, rgb); ref.build("vcs")
                ref.trace_rays(rays[:65536], algo, want_hits=False)                       # warm-up
                r = ref.trace_rays(rays, algo, want_hits=False)
                rms = float(ref.lib.refg_last_kernel_ms())
                row.update(ref_gpu_ms=rms, ref_gpu_mrays_per_s=n / rms / 1e3, speedup_vs_ref_gpu=rms / ms)
                got = d_col.cpu().numpy().astype(np.uint32)
                row.update(colour_equal_vs_ref_gpu_fmad=float((got == r["colour"]).mean()))
                ref.close()
            if po.available("refh"):
                po.set_lighting("refh")
                refh = po.OracleScene("refh"); refh.add_voxels(xyz, rgb); refh.build("vcs")
                sub = 1 << 20
                t0 = time.perf_counter()
                want = refh.trace_rays(rays[:sub], algo, threads=os.cpu_count() or 1)
                cpu_s = time.perf_counter() - t0
                got = s.trace_rays(rays[:sub], algo, want_hits=True)
                row.update(hit_parity_first_1M=float((got["hits"] == want["hits"]).all(-1).mean()), colour_parity_first_1M=float((got["colour"] == want["colour"]).mean()),
                           ref_cpu_mrays_per_s=sub / cpu_s / 1e6, ref_cpu_cores=os.cpu_count() or 1)
                refh.close()
            print(json.dumps(row), flush=True)
            out.append(row)
        s.close()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
