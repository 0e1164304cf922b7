#!/bin/bash
# N = 8, strong-scaling leg only: where the next view is claimed (VRM_CLAIM_TAIL: views left below which a rank claims only when it is free)
cd "$(dirname "$0")/.."
for t in 0 16 64; do
  VRM_CLAIM_TAIL=$t python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --orbit-only 2> gpurun_out/n8_orbit_tail$t.err | tail -1 > gpurun_out/n8_orbit_tail$t.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/n8_orbit_tail$t.json"))["orbit_2048_strong_scaling"]
    print("tail $t", {k: (round(d[k]["ms_per_orbit"], 2), d[k]["views_per_rank"]) for k in ("longestaxis", "original")}, d.get("error"))
except Exception as e:
    print("tail $t failed", e)
PY
done
