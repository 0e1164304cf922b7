"""L2 roofs of the traversal kernels on this GPU: 8-byte random gathers and coalesced streaming reads from
an L2-resident working set, for several working-set sizes (the VCS headers the 4K terrain frame touches are ~10 MB; the
126 MB L2 holds every size below).  Writes gpurun_out/l2_bw.json; the numbers are recorded in BASELINE.md 3 and used by bench.py.

    python tools/l2_bw.py [--out gpurun_out/l2_bw.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelraymarcher_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/l2_bw.json")
    a = ap.parse_args()
    rows = []
    for mb in (1, 4, 16, 32, 64, 96, 256, 1024):
        best = None
        for _ in range(3):
            r = api.microbench_l2(0, mb << 20)
            if best is None or r["gather8_loads_per_ns"] > best["gather8_loads_per_ns"]:
                best = r
        best["working_set_mb"] = mb
        rows.append(best)
        print(json.dumps(best), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
