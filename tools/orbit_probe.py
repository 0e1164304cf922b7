"""Per-view kernel times of the config-4 orbit (2048^3 sparse shells generated on the GPU, 64 views, 1920x1080) on a GPU box:
shows which views carry the reference's epsilon-crawl pathologies and what the other views cost."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelraymarcher_b200 import api

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
w, h = 1920, 1080
s = api.VoxelScene(0)
t0 = time.perf_counter(); n = s.generate_sparse_shells(size, 64, 7, 35); t1 = time.perf_counter()
ms = s.generate_voxel_scene("vcs")
print(f"{n} voxels generated on the GPU in {1e3 * (t1 - t0):.1f} ms, built in {ms:.1f} ms, info {s.info()}", flush=True)
rows = []
for algo in ("longestaxis", "original"):
    times = []
    for v in range(64):
        ang = 2.0 * np.pi * (v + 0.37) / 64
        r, el = 1.5 * size / 2, np.deg2rad(20.0)
        c = size / 2
        org = (float(c + r * np.cos(el) * np.cos(ang)), float(c + r * np.sin(el)), float(c + r * np.cos(el) * np.sin(ang)))
        cam = api.Camera(org, (c, c, c), (0.0, 1.0, 0.0), 60.0, np.float32(w) / np.float32(h))
        s.render(w, h, algo, cam)
        times.append(s.render(w, h, algo, cam)["kernel_ms"])
    t = np.array(times)
    order = np.argsort(-t)
    print(f"{algo}: sum {t.sum():.1f} ms, median {np.median(t):.3f} ms, slowest views {[(int(i), round(float(t[i]), 2)) for i in order[:6]]}", flush=True)
    rows.append(dict(algo=algo, per_view_ms=[float(x) for x in t]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/orbit_probe_{size}.json", "w"))
