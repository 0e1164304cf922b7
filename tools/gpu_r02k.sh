#!/bin/bash
# Round 2, GPU call k: fused form with 9 CTAs per SM + per-warp stores for local frames; A/B of per-warp stores in the queue pipelines; bench line.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02k_pytest.log
tail -6 gpurun_out/r02k_pytest.log
timeout 1200 python tools/ab_variants.py main wsq main@vcs:longestaxis c9w@vcs:longestaxis > gpurun_out/r02k_ab.log 2>&1; cat gpurun_out/r02k_ab.log
cp gpurun_out/ab.json gpurun_out/r02k_ab.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err || { echo "bench failed"; tail -20 gpurun_out/r02k_bench_n1.err; }
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02k_bench_n1.json"))
    for k in ("value", "ms_per_step", "e2e", "single_view", "build", "build_hashtable", "gpu_launches"):
        print(k, json.dumps(d.get(k))[:400])
    print("combos", json.dumps({k: (round(v["ms_per_frame"], 3), round(v.get("speedup_vs_ref_gpu", 0), 1)) for k, v in d.get("combos", {}).items()}))
    print("orbit", json.dumps(d.get("orbit_2048_strong_scaling"))[:600])
except Exception as e:
    print("no bench line", e)
PY
du -sh gpurun_out
