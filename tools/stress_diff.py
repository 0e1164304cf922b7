"""CPU differential stress of the PRODUCT's traversal source (tests/hostsim: vrm_core.cuh nested form + vrm_flat.cuh state machine)
against the C oracle: random rays from pseudo-random origins, many of them on integer coordinates, cluster faces (multiples of 8)
and region faces (multiples of 64), compared in colour, hit voxel AND event counters.  This is the tool that found the two
crawl_skip defects of round 1; run it after touching any fast-forward.

    python tools/stress_diff.py [seed] [rays-per-origin]

Oracle = oracle/vrm_oracle.c ("orc"): it has defined behaviour where the reference's host build reads past its 512-entry cluster
table (a ray rebased onto local coordinate 64.0 -- origins with a coordinate on a region face provoke it; such mismatches are reported as
"origin on a region face" and are not counted as defects: nothing defined exists to match there -- the oracle's own event counters
vary from run to run for such rays, its cluster table being followed by the next region's struct.  Lookups with y or z = 64 ARE defined
(empty) and exact since round 1; colour / hit differences from such origins are therefore listed separately at the end)."""
import os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import build_oracle, po  # noqa: E402
from voxelraymarcher_b200 import scenes  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    po.set_lighting("orc"); po.set_lighting("sim")
    lib = po._lib("sim")
    rng = np.random.default_rng(seed)
    cases = [("shells512", scenes.sparse_shells(512, 64, seed=21, fill_pct=30), 1), ("terrain192", scenes.terrain(192, 9), 1), ("probe", scenes.probe_scene(), 8)]
    defects = 0
    corner_pixels = 0
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
            a, b = build_oracle("orc", xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
            for oi, org in enumerate(origins):
                corner = any((c * scale) % 64 == 0 for c in org)   # a ray rebased onto local x = 64.0 indexes the reference's cluster table at >= 512
                rays = scenes.random_rays(n_rays, org, seed=300 + oi)
                ta = a.trace_rays(rays, algo, scale=scale, want_counters=True)
                for flat in (1, 0):
                    lib.sim_set_flat(flat)
                    tb = b.trace_rays(rays, algo, scale=scale, want_counters=True)
                    if not all(np.array_equal(ta[k], tb[k]) for k in ("colour", "hits", "counters")):
                        bad = int(((ta["colour"] != tb["colour"]) | (ta["hits"] != tb["hits"]).any(1)).sum())
                        print(f"MISMATCH{' (origin on a region face: the reference can be undefined)' if corner else ''} {name} {storage} {algo} origin {org} {'state machine' if flat else 'nested'}: "
                              f"oracle {ta['counters'][:3]} product {tb['counters'][:3]}, {bad} rays differ in colour / hit", flush=True)
                        defects += 0 if corner else 1
                        corner_pixels += bad if corner else 0
            print(name, storage, algo, "done", flush=True)
            a.close(); b.close()
    lib.sim_set_flat(1)
    print("defects:", defects, " rays differing in colour / hit from region-face origins (undefined corner):", corner_pixels)
    return 1 if defects else 0


if __name__ == "__main__":
    sys.exit(main())
