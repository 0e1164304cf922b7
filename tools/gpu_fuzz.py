"""The differential stress of tools/stress_diff.py through the real kernels (C ABI, trace_rays with statistics) on a GPU box:
random rays from grid-aligned origins, all four combinations, colours + hit voxels + event counters against the C oracle.

    python tools/gpu_fuzz.py [seed] [rays-per-origin]
"""
import os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import build_oracle, po  # noqa: E402
from voxelraymarcher_b200 import api, scenes  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
    po.set_lighting("orc")
    rng = np.random.default_rng(seed)
    cases = [("shells512", scenes.sparse_shells(512, 64, seed=21, fill_pct=30), 1), ("terrain192", scenes.terrain(192, 9), 1), ("probe", scenes.probe_scene(), 8)]
    defects = 0
    for name, (xyz, rgb), scale in cases:
        lo, hi = xyz.min(0), xyz.max(0)
        origins = []
        for _ in range(10):
            p = rng.integers(lo - 30, hi + 30, 3).astype(np.float64)
            r = rng.random()
            if r < 0.35:
                p = np.round(p / 64) * 64
            elif r < 0.7:
                p = np.round(p / 8) * 8
            if rng.random() < 0.3:
                p = p + rng.choice([0.5, 0.25, 0.125])
            origins.append(tuple((p / scale).tolist()))
        for storage, algo in (("vcs", "longestaxis"), ("vcs", "original"), ("hashtable", "longestaxis"), ("hashtable", "original")):
            a = build_oracle("orc", xyz, rgb, storage)
            s = api.VoxelScene(0)
            s.add_voxels(xyz, rgb)
            s.generate_voxel_scene(storage)
            s.set_statistics(True)
            for oi, org in enumerate(origins):
                corner = any((c * scale) % 64 == 0 for c in org)   # a ray rebased onto local x = 64.0 indexes the reference's cluster table at >= 512
                rays = scenes.random_rays(n_rays, org, seed=300 + oi)
                want = a.trace_rays(rays, algo, scale=scale, want_counters=True)
                got = s.trace_rays(rays, algo, scale=scale, want_hits=True)
                st = s.get_statistics()
                cnt = (st["exist_checks"], st["exist_false"], st["lookups"])
                ok = np.array_equal(got["colour"], want["colour"]) and np.array_equal(got["hits"], want["hits"]) and cnt == tuple(int(v) for v in want["counters"][:3])
                if not ok:
                    bad = int(((got["colour"] != want["colour"]) | (got["hits"] != want["hits"]).any(1)).sum())
                    print(f"MISMATCH{' (origin on a region face: the reference can be undefined)' if corner else ''} {name} {storage} {algo} origin {org}: oracle {want['counters'][:3]} kernels {cnt}, {bad} rays differ", flush=True)
                    defects += 0 if corner else 1
            print(name, storage, algo, "done", flush=True)
            a.close(); s.close()
    print("defects:", defects)
    return 1 if defects else 0


if __name__ == "__main__":
    sys.exit(main())
