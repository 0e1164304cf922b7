#!/bin/bash
# Round-end measurement set on a GPU box (one GPU): whole GPU test tier, captures of the final kernels (summarised on the box, merged into
# profiles/ncu_summary.json BEFORE the bench so that its issue roof uses this build's capture), bench line + reference arm, launch list, A/B
# against voxelraymarcher_b200/variants/libvrm_<BASE>.so when that library travelled with the snapshot.
#   bash tools/profile_round.sh <tag> [<base variant>] [full]     ("full": captures of every kernel, else the headline + the changed ones)
tag=${1:-r02z}; base=${2:-prev}; full=${3:-}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
for m in 0 3 4; do
	( VRM_RENDER_MODE=$m timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) > gpurun_out/${tag}_pytest_mode$m.log 2>&1; echo "mode $m: $(tail -1 gpurun_out/${tag}_pytest_mode$m.log)"
done
rm -f gpurun_out/ncu_summary.json
bash tools/gpu_capture.sh ${tag}_ncu_render_vcs_longestaxis render_kernel 3 terrain512_4k_vcs_longestaxis -- python bench.py --steps 2 --warmup 3 --no-baselines --single-view
bash tools/gpu_capture.sh ${tag}_ncu_render_hashtable_longestaxis render_kernel 6 terrain512_4k_hashtable_longestaxis -- python tools/explore.py --iters 2 --combos hashtable:longestaxis --out gpurun_out/x.json
bash tools/gpu_capture.sh ${tag}_ncu_trace_config5_fused trace_fused_kernel 2 shells1024_trace_vcs_longestaxis -- python tools/ncu_targets.py trace5 longestaxis
bash tools/gpu_capture.sh ${tag}_ncu_orbit_config4_vcs_longestaxis render_kernel 2 shells2048_orbit_vcs_longestaxis -- python tools/ncu_targets.py orbit4
if [ "$full" = "full" ]; then
	bash tools/gpu_capture.sh ${tag}_ncu_render_hashtable_original render_kernel 6 terrain512_4k_hashtable_original -- python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json
	bash tools/gpu_capture.sh ${tag}_ncu_shadow_hashtable_original shadow_kernel 6 terrain512_4k_hashtable_original_shadow -- python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json
	bash tools/gpu_capture.sh ${tag}_ncu_render_vcs_original render_kernel 6 terrain512_4k_vcs_original -- python tools/explore.py --iters 2 --combos vcs:original --out gpurun_out/x.json
	bash tools/gpu_capture.sh ${tag}_ncu_shadow_vcs_original shadow_kernel 6 terrain512_4k_vcs_original_shadow -- python tools/explore.py --iters 2 --combos vcs:original --out gpurun_out/x.json
fi
rm -f gpurun_out/x.json
python - <<'PY'
import json, os
new = json.load(open("gpurun_out/ncu_summary.json")) if os.path.exists("gpurun_out/ncu_summary.json") else {}
path = "profiles/ncu_summary.json"
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(new)
json.dump(old, open(path, "w"), indent=1)
print("ncu_summary workloads updated:", sorted(new))
PY
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err || { echo "bench failed"; tail -20 gpurun_out/${tag}_bench_n1.err; }
python - "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench_n1.json"))
    for k in ("value", "ms_per_step", "e2e", "single_view", "stats_per_ray", "build", "build_hashtable", "cpu_baseline", "gpu_launches"):
        print(k, json.dumps(d.get(k))[:500])
    print("combos", json.dumps({k: (round(v["ms_per_frame"], 3), round(v.get("speedup_vs_ref_gpu", 0), 1)) for k, v in d.get("combos", {}).items()}))
    print("orbit", json.dumps(d.get("orbit_2048_strong_scaling"))[:500])
    r = d["roofline"]; print("roofline", json.dumps({k: r[k] for k in ("achieved", "frac", "bytes_per_ray", "issue", "l2", "traffic")}))
except Exception as e:
    print("no bench line", e)
PY
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; cut -c1-330 gpurun_out/${tag}_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-baselines > gpurun_out/${tag}_ncu_launches_bench.log 2>&1
if [ -f voxelraymarcher_b200/variants/libvrm_$base.so ]; then
	timeout 900 python tools/ab_variants.py $base main > gpurun_out/${tag}_ab.log 2>&1; cat gpurun_out/${tag}_ab.log
	cp gpurun_out/ab.json gpurun_out/${tag}_ab.json
fi
timeout 300 python tools/ncu_targets.py trace5 longestaxis 2>&1 | tail -2
timeout 300 python tools/ncu_targets.py trace5 original 2>&1 | tail -2
du -sh gpurun_out
