#!/bin/bash
# One `ncu --set full` capture on the GPU box, summarised THERE (gpurun only brings back 64 MiB, a report with per-instruction counters
# is ~15 MB): writes gpurun_out/<tag>.json, <tag>_opcodes.txt (tools/ncu_summarize.py), <tag>_sass.csv.gz (the source page: per SASS
# instruction executed / thread-executed / samples / stall reasons) and keeps the .ncu-rep only when KEEP_REP=1.
#   bash tools/gpu_capture.sh <tag> <kernel regex> <launches to skip> [workload key or -] -- <command ...>
tag=$1; regex=$2; skip=$3; key=$4; shift 5
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$regex" -s "$skip" -c 1 -o gpurun_out/$tag -f "$@" > gpurun_out/${tag}_ncu.log 2>&1
if [ ! -f gpurun_out/$tag.ncu-rep ]; then echo "$tag: no report"; tail -5 gpurun_out/${tag}_ncu.log; exit 0; fi
if [ "$key" = "-" ]; then key=""; fi
python tools/ncu_summarize.py gpurun_out/$tag.ncu-rep gpurun_out/$tag $key > gpurun_out/${tag}_summary.log 2>&1 || tail -3 gpurun_out/${tag}_summary.log
ncu -i gpurun_out/$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/${tag}_sass.csv.gz
ncu -i gpurun_out/$tag.ncu-rep --page details --csv 2>/dev/null | gzip -9 > gpurun_out/${tag}_details.csv.gz
[ "$KEEP_REP" = "1" ] || rm -f gpurun_out/$tag.ncu-rep
python - <<PY
import json
d = json.load(open("gpurun_out/$tag.json"))["launches"][0]
print("$tag:", {k: d.get(k) for k in ("kernel", "duration_ms", "warp_instructions", "threads_per_instruction", "issue_slot_utilisation_pct", "achieved_occupancy_pct", "dram_bytes", "l2_hit_pct", "l1_hit_pct", "registers_per_thread")})
PY
