#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: primary Mrays/s and ms/frame at 4K on the 512^3
scene at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's own CPU implementation on the host cores)

A step = one pass of the hot path over one batch of synthetic input = ONE 3840x2160 frame per GPU of the procedural
512^3 terrain (~32 M voxels, BASELINE.json configs[2]: "single view") with the reference's default CLI combination
(Voxel Cluster Store + longest-axis traversal, Main.cu:45-68), SURVEY.md 8d-3's camera, shadows on.  The voxel structure is built once per GPU, on the GPU, before the timed region and is replicated
on every rank; at N>1 every step ends with the NCCL gather of the finished frames on rank 0 (the path's only
exchange step), so `value` includes it.

value  = W*H*N*K / t, t = max over ranks of the summed per-step CUDA-event time (render kernel [+ gather]); inputs
         and outputs resident in HBM.  L2 is flushed between steps (outside the per-step event pairs).
e2e    = the same metric through the host-buffer C-ABI call vrm_render (camera H2D + frame D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 3840, 2160
SCENE_SIZE, SCENE_SEED = 512, 1234
STORAGE, ALGORITHM = "vcs", "longestaxis"       # the reference CLI's defaults
ORBIT_CENTRE = (256.0, 64.0, 256.0)
ORBIT_RADIUS, ORBIT_HEIGHT, ORBIT_VIEWS = 498.0, 352.0, 64
METRIC = "primary_mrays_per_s_4k_512cube"
UNIT = "Mrays/s"


def orbit_camera(api, view: int):
    """View `view` of a 64-view orbit around the terrain; view 0 is SURVEY.md §8d-3's camera (-96,352,-96)."""
    ang = np.float64(-0.75 * np.pi) + 2.0 * np.pi * (view % ORBIT_VIEWS) / ORBIT_VIEWS
    org = (float(np.float32(ORBIT_CENTRE[0] + ORBIT_RADIUS * np.cos(ang))), ORBIT_HEIGHT, float(np.float32(ORBIT_CENTRE[2] + ORBIT_RADIUS * np.sin(ang))))
    return api.Camera(org, ORBIT_CENTRE, (0.0, 1.0, 0.0), 60.0, np.float32(WIDTH) / np.float32(HEIGHT))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(mx)) if mx else None, reasons=sorted(reasons), samples=len(sm))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(stats, storage):
    """SURVEY.md §8d per-ray model summed over the frame: hashtable 4*P1 + 4*P2 + 4*H; VCS 4*E + 4*L + 4*S + 4*H with S = 0
    (occupancy mask + popcount rank replaces the binary search); + 4 B per region-table entry read (int32 here, 8-byte
    pointers in the reference) + 3 B framebuffer write per pixel."""
    if storage == "hashtable":
        b = 4 * stats["lookups"] + 4 * stats["table2_probes"] + 4 * stats["lookup_hits"]
    else:
        b = 4 * stats["exist_checks"] + 4 * stats["lookups"] + 4 * stats["lookup_hits"]
    return b + 4 * stats["region_reads"] + 3 * stats["rays"]


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this workload (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        return json.load(open(path)).get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref host build of the unmodified
    reference; the C restatement only if that build is absent) on all host threads, same scene / camera / metric."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    from voxelraymarcher_b200 import scenes
    kind = "refh" if po.available("refh") else "orc"
    cores = os.cpu_count() or 1
    xyz, rgb = scenes.terrain(SCENE_SIZE, SCENE_SEED)
    po.set_lighting(kind)
    ref = po.OracleScene(kind)
    ref.add_voxels(xyz, rgb)
    ref.build(STORAGE)
    cams = [po.make_camera(*_orbit_args(0), kind) for v in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        ref.render(cams[i], WIDTH, HEIGHT, ALGORITHM, want_hits=False, threads=cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        ref.render(cams[args.warmup + i], WIDTH, HEIGHT, ALGORITHM, want_hits=False, threads=cores)
    dt = time.perf_counter() - t0
    value = WIDTH * HEIGHT * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "refh" else "port",
                         "sample": f"full {WIDTH}x{HEIGHT} frame per step, {args.steps} steps, traversal only (scene build excluded)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _orbit_args(view):
    ang = np.float64(-0.75 * np.pi) + 2.0 * np.pi * (view % ORBIT_VIEWS) / ORBIT_VIEWS
    org = (float(np.float32(ORBIT_CENTRE[0] + ORBIT_RADIUS * np.cos(ang))), ORBIT_HEIGHT, float(np.float32(ORBIT_CENTRE[2] + ORBIT_RADIUS * np.sin(ang))))
    return org, ORBIT_CENTRE, (0.0, 1.0, 0.0), 60.0, np.float32(WIDTH) / np.float32(HEIGHT)


def workload_config():
    return {"workload": f"terrain{SCENE_SIZE}_4k_{STORAGE}_{ALGORITHM}", "scene": f"procedural {SCENE_SIZE}^3 terrain, seed {SCENE_SEED}, ~32 M voxels (BASELINE.json configs[2])",
            "resolution": f"{WIDTH}x{HEIGHT}", "storage": STORAGE, "algorithm": ALGORITHM, "shadows": True,
            "views": "one frame per GPU per step, camera (-96,352,-96) -> (256,64,256), fov 60 (SURVEY.md 8d-3)", "l2": "flushed between steps (512 MiB write, outside the per-step event pairs)",
            "parallelism": "frames sharded across GPUs (one per GPU per step), structure replicated, NCCL gather of frames to rank 0"}


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the cpu_baseline / ref_gpu legs (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on stdout, so everything
    # before the final print goes to stderr.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from voxelraymarcher_b200 import api, scenes

    if not torch.cuda.is_available() or not api.device_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    # ---- scene: generated on the host (integer-only, deterministic), built on the GPU, replicated per rank --------
    xyz, rgb = scenes.terrain(SCENE_SIZE, SCENE_SEED)
    warm = api.VoxelScene(local_rank)            # loads the CUDA modules so that the timed build below is not a cold start
    warm.add_voxels(xyz[:65536], rgb[:65536])
    warm.generate_voxel_scene(STORAGE)
    warm.close()
    scene = api.VoxelScene(local_rank)
    scene.add_voxels(xyz, rgb)
    build_ms = scene.generate_voxel_scene(STORAGE)
    info = scene.info()
    stream = torch.cuda.Stream(dev)          # a dedicated non-default stream: kernels, flushes, NCCL ordering and events all live here
    torch.cuda.set_stream(stream)
    scene.set_stream(stream.cuda_stream)

    frame = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.uint8, device=dev)
    # Exchange step (N > 1): the finished frames are gathered on rank 0.  Preferred form: FUSED into the render kernel --
    # every rank maps rank 0's gather buffer through CUDA IPC and its kernel stores the frame straight into its slot over
    # NVLink / NVSwitch (coalesced 96-byte row segments); a 4-byte NCCL all-reduce per step tells rank 0 the slots are
    # complete.  Fallback (no peer access): render locally, then an NCCL gather of the 24.9 MB frames.
    peer, exchange = None, "none"
    if world > 1:
        from voxelraymarcher_b200 import multigpu
        try:
            peer = multigpu.PeerFrameBuffer(world, WIDTH, HEIGHT, local_rank)
            exchange = "fused: render kernel stores into rank 0's buffer over NVLink peer memory (CUDA IPC) + 4-byte NCCL all-reduce per step"
        except Exception as exc:  # noqa: BLE001
            peer, exchange = None, f"nccl gather of frames (peer mapping failed: {exc})"
        flags = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if int(flags.item()) == 0 and peer is not None:
            peer.close()
            peer, exchange = None, "nccl gather of frames (peer mapping failed on another rank)"
    out_ptr = peer.ptr_for(rank) if peer is not None else frame.data_ptr()
    token = torch.zeros(1, dtype=torch.int32, device=dev)
    gathered = [torch.zeros_like(frame) for _ in range(world)] if (world > 1 and rank == 0 and peer is None) else None
    flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
    total = args.warmup + args.steps
    cams = [orbit_camera(api, 0) for i in range(total)]   # configs[2] is a single view: every rank renders SURVEY.md §8d-3's camera

    def step(i, ev0, ev1, evk):
        flush.add_(1)                       # evict L2 (512 MiB > 126 MB), not timed
        ev0.record(stream)
        scene.render_device(WIDTH, HEIGHT, ALGORITHM, cams[i], out_ptr)
        evk.record(stream)
        if world > 1:
            if peer is not None:
                dist.all_reduce(token)           # completion signal: after it, every rank's frame is in rank 0's buffer
            else:
                dist.gather(frame, gathered, dst=0)
        ev1.record(stream)

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(total)]
    for i in range(args.warmup):
        step(i, *evs[i])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)                          # let nvidia-smi come up: the timed region itself is only tens of milliseconds
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    for i in range(args.warmup, total):
        step(i, *evs[i])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    step_ms = [evs[i][0].elapsed_time(evs[i][1]) for i in range(args.warmup, total)]
    kern_ms = [evs[i][0].elapsed_time(evs[i][2]) for i in range(args.warmup, total)]
    t_ms = float(sum(step_ms))
    tk_ms = float(sum(kern_ms))

    # ---- e2e: host-buffer C-ABI call (camera H2D + frame D2H inside the timed region), pinned result buffer ---------
    host_frame = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.uint8).pin_memory()
    host_np = host_frame.numpy()
    e2e_s = 0.0
    for i in range(total):
        flush.add_(1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        scene.render(WIDTH, HEIGHT, ALGORITHM, cams[i], rgb_out=host_np)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e_s += dt
    clocks = sampler.stop() if rank == 0 else None

    exchange_ok = None
    if peer is not None:
        # outside the timed region: the gathered slots must equal a local render of the same camera
        scene.render_device(WIDTH, HEIGHT, ALGORITHM, cams[-1], frame.data_ptr())
        torch.cuda.synchronize(dev)
        dist.barrier()
        if rank == 0:
            full = peer.to_tensor()
            exchange_ok = bool(all(torch.equal(full[r], frame) for r in range(world)))
    if world > 1:
        red = torch.tensor([t_ms, tk_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        t_ms, tk_ms, e2e_s = [float(v) for v in red.tolist()]

    rays_total = WIDTH * HEIGHT * args.steps * n_gpus
    value = rays_total / (t_ms * 1e-3) / 1e6
    value_no_gather = rays_total / (tk_ms * 1e-3) / 1e6
    e2e_value = rays_total / e2e_s / 1e6

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (render): algorithmic bytes from the kernel's own event counters ------
        scene.set_statistics(True)
        scene.render_device(WIDTH, HEIGHT, ALGORITHM, cams[args.warmup], frame.data_ptr())
        scene.synchronize()
        stats = scene.get_statistics()
        scene.set_statistics(False)
        peak, peak_src = measured_peak()
        k_ms = float(np.mean(kern_ms))
        alg_bytes = algorithmic_bytes(stats, STORAGE)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        cfg = workload_config()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(cfg["workload"]),
                    "kernel": "render_kernel<VCS,LongestAxis,flat-loop>", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                    "bytes_per_ray": alg_bytes / stats["rays"], "peak_source": peak_src,
                    "note": "latency/issue-bound gather walk: the touched working set is L2-resident, so HBM traffic is far below the algorithmic bytes (see DESIGN.md roofline)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 60, "d2h_bytes_per_step": WIDTH * HEIGHT * 3},
            "gpu_launches": 2 * args.steps,   # per step: render_kernel + the one-block resume_kernel behind it (parked rays; ~3 us)
            "roofline": roofline, "exchange": exchange, "exchange_verified": exchange_ok,
            "ms_per_frame_kernel": tk_ms / args.steps, "value_without_gather": value_no_gather, "wall_s_timed_region": wall,
            "build": {"ms": build_ms, "mvoxels_per_s": xyz.shape[0] / build_ms / 1e3, "voxels": int(xyz.shape[0]), "unique_voxels": info["unique_voxels"],
                      "regions": info["filled"], "structure_bytes": info["bytes"]},
            "stats_per_ray": {k: stats[k] / stats["rays"] for k in ("exist_checks", "exist_false", "lookups", "lookup_hits", "region_reads")},
        }
        if n_gpus == 1 and not args.no_baselines:
            line.update(baseline_legs(api, scene, xyz, rgb, cams[args.warmup], flush, dev))
    scene.close()
    if world > 1:
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)


def baseline_legs(api, scene, xyz, rgb, cam, flush, dev):
    """Baselines timed in the same run (north star): the reference's own CUDA kernels rebuilt for sm_100a on this GPU, and
    the reference's traversal built for the host cores.  Reported, never the thing measured.  Also times the other three
    storage x algorithm combinations of the native path on the same frame."""
    import torch
    from oracle import pyoracle as po
    out = {}
    cores = os.cpu_count() or 1
    # native: all four combinations, median of 7 launches, L2 flushed
    combos = {}
    scenes_by_storage = {STORAGE: scene}
    for storage in ("vcs", "hashtable"):
        if storage not in scenes_by_storage:
            warm = api.VoxelScene(scene.device)      # first use of this storage's builder kernels / allocations: not part of the timed build
            warm.add_voxels(xyz[:65536], rgb[:65536])
            warm.generate_voxel_scene(storage)
            warm.close()
            s = api.VoxelScene(scene.device)
            s.add_voxels(xyz, rgb)
            s.generate_voxel_scene(storage)
            scenes_by_storage[storage] = s
        s = scenes_by_storage[storage]
        fb = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.uint8, device=dev)
        cur = torch.cuda.current_stream(dev)
        s.set_stream(cur.cuda_stream)
        for algo in ("longestaxis", "original"):
            times = []
            for i in range(9):
                flush.add_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur); s.render_device(WIDTH, HEIGHT, algo, cam, fb.data_ptr()); e1.record(cur)
                torch.cuda.synchronize(dev)
                if i >= 2:
                    times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            combos[f"{storage}+{algo}"] = {"ms_per_frame": ms, "mrays_per_s": WIDTH * HEIGHT / ms / 1e3, "build_ms": s.build_ms}
    # reference CUDA kernels (default nvcc flags) on the same frame
    if po.available("refg"):
        po.set_lighting("refg")
        for storage in ("vcs", "hashtable"):
            ref = po.OracleScene("refg")
            t0 = time.perf_counter()
            ref.add_voxels(xyz, rgb)
            ref.build(storage)
            host_build_s = time.perf_counter() - t0
            for algo in ("longestaxis", "original"):
                ms = np.zeros(5, np.float32)
                rc = ref.lib.refg_render_timed(ref.h, po._ptr(cam.data), po._ptr(np.zeros(3, np.float32)), 1, po.ALGORITHM[algo], WIDTH, HEIGHT, 2, 5, po._ptr(ms))
                if rc == 0:
                    m = float(np.median(ms))
                    c = combos[f"{storage}+{algo}"]
                    c.update(ref_gpu_ms_per_frame=m, ref_gpu_mrays_per_s=WIDTH * HEIGHT / m / 1e3, speedup_vs_ref_gpu=m / c["ms_per_frame"], ref_host_build_s=host_build_s)
            ref.close()
    out["combos"] = combos
    # reference traversal on the host cores (cpu_baseline)
    kind = "refh" if po.available("refh") else "orc"
    po.set_lighting(kind)
    ref = po.OracleScene(kind)
    ref.add_voxels(xyz, rgb)
    ref.build(STORAGE)
    ref.render(cam.data, WIDTH, HEIGHT, ALGORITHM, want_hits=False, threads=cores)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ref.render(cam.data, WIDTH, HEIGHT, ALGORITHM, want_hits=False, threads=cores)
    dt = time.perf_counter() - t0
    ref.close()
    out["cpu_baseline"] = {"value": WIDTH * HEIGHT * reps / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "refh" else "port",
                           "sample": f"{reps} full {WIDTH}x{HEIGHT} frames of the same scene/camera, traversal only (scene build excluded)"}
    for s in scenes_by_storage.values():
        if s is not scene:
            s.close()
    return out


if __name__ == "__main__":
    main()
