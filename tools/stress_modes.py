"""Two more differential stress modes for the traversal source (tests/hostsim) against the C oracle, next to tools/stress_diff.py:

    python tools/stress_modes.py ulps [seed]   origins on region / cluster faces and integers, each coordinate nudged by -3..+3 ulps
    python tools/stress_modes.py dirs [seed]   ordinary origins; one or two direction components shrunk by 1e-3 .. 1e-12 (not renormalised)

Both compare colour, hit voxel and event counters for all four combinations and both forms (state machine, nested).  Round 1: no
mismatch in `ulps` (seeds 1, 2) nor in the part of `dirs` seed 1 that finished (terrain: all four combinations; shells: VCS + longest
axis) -- `dirs` is slow where the oracle has to execute the EPSILON crawls step by step (hours for shells + original)."""
import os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import build_oracle, po  # noqa: E402
from voxelraymarcher_b200 import scenes  # noqa: E402


def nudge(v, k):
    x = np.float32(v)
    for _ in range(abs(k)):
        x = np.nextafter(x, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
    return float(x)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "ulps"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    po.set_lighting("orc"); po.set_lighting("sim")
    lib = po._lib("sim")
    rng = np.random.default_rng(seed)
    cases = [("terrain192", scenes.terrain(192, 9), 1), ("shells256", scenes.sparse_shells(256, 32, seed=9, fill_pct=45), 1)]
    defects = 0
    for name, (xyz, rgb), scale in cases:
        lo, hi = xyz.min(0), xyz.max(0)
        origins = []
        for _ in range(8 if mode == "ulps" else 6):
            p = rng.integers(lo, hi, 3).astype(np.float64)
            if mode == "ulps":
                grid = rng.choice([64, 8, 1])
                p = np.round(p / grid) * grid
                p = [nudge(c, int(rng.integers(-3, 4))) if rng.random() < 0.7 else float(c) for c in p]
                origins.append(tuple(p))
            else:
                if rng.random() < 0.4:
                    p = np.round(p / 8) * 8
                if rng.random() < 0.5:
                    p = p + rng.random(3)
                p = np.where(p % 64 == 0, p + 1, p)   # keep off region faces (the reference's undefined corner)
                origins.append(tuple(p.tolist()))
        for storage, algo in (("vcs", "longestaxis"), ("vcs", "original"), ("hashtable", "longestaxis"), ("hashtable", "original")):
            a, b = build_oracle("orc", xyz, rgb, storage), build_oracle("sim", xyz, rgb, storage)
            for oi, org in enumerate(origins):
                rays = scenes.random_rays(15000 if mode == "ulps" else 8000, org, seed=500 + oi + 10 * seed).copy()
                if mode == "dirs":
                    k = rng.integers(0, 3, len(rays))
                    e = (10.0 ** -rng.integers(3, 13, len(rays))).astype(np.float32)
                    idx = np.nonzero(rng.random(len(rays)) < 0.7)[0]
                    rays[idx, 3 + k[idx]] *= e[idx]
                    idx2 = np.nonzero(rng.random(len(rays)) < 0.2)[0]
                    rays[idx2, 3 + (k[idx2] + 1) % 3] *= np.float32(1e-5)
                ta = a.trace_rays(rays, algo, scale=scale, want_counters=True)
                for flat in (1, 0):
                    lib.sim_set_flat(flat)
                    tb = b.trace_rays(rays, algo, scale=scale, want_counters=True)
                    bad = int(((ta["colour"] != tb["colour"]) | (ta["hits"] != tb["hits"]).any(1)).sum())
                    if bad or not np.array_equal(ta["counters"], tb["counters"]):
                        corner = mode == "ulps" and any(round(c) % 64 == 0 for c in org)
                        print(f"MISMATCH{' (origin at a region face +- ulps)' if corner else ''}", name, storage, algo, org, "state machine" if flat else "nested", bad,
                              ta["counters"][:3], tb["counters"][:3], flush=True)
                        defects += 0 if corner else 1
            print(name, storage, algo, "done", flush=True)
            a.close(); b.close()
    lib.sim_set_flat(1)
    print("defects:", defects)
    return 1 if defects else 0


if __name__ == "__main__":
    sys.exit(main())
