#!/bin/bash
# Round 2, third GPU call: the shadow-ray queue -- whole GPU test tier, A/B against the previous build and the tuning variants,
# configs 4 / 5 timings, bench line, captures of the primary and the shadow kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02d_pytest.log
tail -16 gpurun_out/r02d_pytest.log
timeout 1500 python tools/ab_variants.py main prev hv16 hv8 rv16 wstore smb3 smb6 > gpurun_out/r02d_ab.log 2>&1; cat gpurun_out/r02d_ab.log
cp gpurun_out/ab.json gpurun_out/r02d_ab.json
for lib in main prev; do
  if [ $lib = prev ]; then export VRM_B200_LIB=$PWD/voxelraymarcher_b200/variants/libvrm_prev.so; else unset VRM_B200_LIB; fi
  for a in longestaxis original; do
    timeout 300 python tools/ncu_targets.py trace5 $a 2>&1 | grep trace5 | sed "s/^/$lib: /"
    timeout 300 python tools/ncu_targets.py orbit4 $a 2>&1 | grep "orbit4" | sed "s/^/$lib: /"
  done
done > gpurun_out/r02d_cfg45.log 2>&1
unset VRM_B200_LIB
cat gpurun_out/r02d_cfg45.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err || { echo "bench failed"; tail -20 gpurun_out/r02d_bench_n1.err; }
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02d_bench_n1.json"))
    for k in ("value", "ms_per_step", "e2e", "single_view", "kernel_ms_per_step", "stats_per_ray"):
        print(k, json.dumps(d.get(k)))
    print("combos", json.dumps({k: round(v["ms_per_frame"], 3) for k, v in d.get("combos", {}).items()}))
    print("orbit", json.dumps(d.get("orbit_2048_strong_scaling")))
except Exception as e:
    print("no bench line", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02d_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-baselines --single-view > gpurun_out/r02d_ncu_launches_bench.log 2>&1
grep -E "render_kernel|shadow_kernel|resume_kernel" gpurun_out/r02d_launches_bench.csv | tail -6 | cut -c1-220
bash tools/gpu_capture.sh r02d_ncu_render_vcs_longestaxis render_kernel 3 - -- python bench.py --steps 2 --warmup 3 --no-baselines --single-view
bash tools/gpu_capture.sh r02d_ncu_shadow_vcs_longestaxis shadow_kernel 3 - -- python bench.py --steps 2 --warmup 3 --no-baselines --single-view
bash tools/gpu_capture.sh r02d_ncu_render_hashtable_original render_kernel 2 - -- python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json
bash tools/gpu_capture.sh r02d_ncu_shadow_hashtable_original shadow_kernel 2 - -- python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json
rm -f gpurun_out/x.json
du -sh gpurun_out
