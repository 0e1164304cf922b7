#!/bin/bash
# Round 2, second GPU call: new multi-GPU entry points (one-GPU tests), the new bench line, launch lists, and the captures VERDICT r01
# asked for (render VCS+LA with source, hash+original, trace_kernel of config 5, a 2048^3 orbit view of config 4, the build kernels).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x ) > gpurun_out/r02c_pytest_multi.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02c_pytest_multi.log
tail -8 gpurun_out/r02c_pytest_multi.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err || { echo "bench failed"; tail -20 gpurun_out/r02c_bench_n1.err; }
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c_bench_n1.json"))
    for k in ("value", "ms_per_step", "e2e", "single_view", "build", "build_hashtable", "orbit_2048_strong_scaling", "kernel_ms_per_step", "views", "clocks"):
        print(k, json.dumps(d.get(k)))
    print("roofline", json.dumps(d["roofline"]))
    print("combos", json.dumps(d.get("combos")))
except Exception as e:
    print("no bench line", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_bench_reference_arm.err; cat gpurun_out/r02c_bench_reference_arm.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-baselines > gpurun_out/r02c_ncu_launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_build_vcs.csv \
    python tools/ncu_targets.py build3 vcs > gpurun_out/r02c_ncu_launches_build_vcs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_build_hash.csv \
    python tools/ncu_targets.py build3 hashtable > gpurun_out/r02c_ncu_launches_build_hash.log 2>&1
KEEP_REP=1 bash tools/gpu_capture.sh r02c_ncu_render_vcs_longestaxis render_kernel 3 terrain512_4k_vcs_longestaxis -- python bench.py --steps 2 --warmup 3 --no-baselines --single-view
bash tools/gpu_capture.sh r02c_ncu_render_hashtable_original render_kernel 2 terrain512_4k_hashtable_original -- python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json
bash tools/gpu_capture.sh r02c_ncu_render_vcs_original render_kernel 2 terrain512_4k_vcs_original -- python tools/explore.py --iters 2 --combos vcs:original --out gpurun_out/x.json
bash tools/gpu_capture.sh r02c_ncu_trace_config5_vcs_longestaxis trace_kernel 2 shells1024_trace_vcs_longestaxis -- python tools/ncu_targets.py trace5 longestaxis
bash tools/gpu_capture.sh r02c_ncu_orbit_config4_vcs_longestaxis render_kernel 2 shells2048_1080p_vcs_longestaxis -- python tools/ncu_targets.py orbit4 longestaxis
for k in radix_scatter radix_hist vcs_fill make_keys; do
  bash tools/gpu_capture.sh r02c_ncu_build_vcs_$k "$k" 3 - -- python tools/ncu_targets.py build3 vcs
done
bash tools/gpu_capture.sh r02c_ncu_build_hash_cuckoo_insert cuckoo_insert 1 - -- python tools/ncu_targets.py build3 hashtable
rm -f gpurun_out/x.json
du -sh gpurun_out; ls gpurun_out | wc -l
