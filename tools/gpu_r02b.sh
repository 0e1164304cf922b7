#!/bin/bash
# Round 2, second GPU call: the lean state machine (VRM_RENDER_MODE=3) -- parity through the GPU tests, A/B against the defaults
# on all four combinations, occupancy variants, config 4/5 timings, one full capture.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
( time VRM_RENDER_MODE=3 timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_configs_gpu.py -m gpu -q -x ) > gpurun_out/r02b_pytest_lean.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest_lean.log
tail -15 gpurun_out/r02b_pytest_lean.log
timeout 1200 python tools/ab_variants.py main main:mode=3 leanmb5:mode=3 leanmb6:mode=3 main:mode=2 > gpurun_out/r02b_ab.log 2>&1; cat gpurun_out/r02b_ab.log
cp gpurun_out/ab.json gpurun_out/r02b_ab.json
for m in 2 3; do
  VRM_RENDER_MODE=$m timeout 300 python tools/ncu_targets.py trace5 longestaxis 2>&1 | grep trace5 | sed "s/^/mode $m: /"
  VRM_RENDER_MODE=$m timeout 300 python tools/ncu_targets.py trace5 original 2>&1 | grep trace5 | sed "s/^/mode $m: /"
  VRM_RENDER_MODE=$m timeout 300 python tools/ncu_targets.py orbit4 longestaxis 2>&1 | grep "orbit4" | sed "s/^/mode $m: /"
  VRM_RENDER_MODE=$m timeout 300 python tools/ncu_targets.py orbit4 original 2>&1 | grep "orbit4" | sed "s/^/mode $m: /"
done > gpurun_out/r02b_cfg45.log 2>&1
cat gpurun_out/r02b_cfg45.log
VRM_RENDER_MODE=3 timeout 900 ncu --set full --import-source on --clock-control none -k regex:render_lean_kernel -s 2 -c 1 -o gpurun_out/prof_r02b_lean_vcs_la -f \
    python tools/explore.py --iters 2 --combos vcs:longestaxis --out gpurun_out/x.json > gpurun_out/r02b_ncu_lean_vcs_la.log 2>&1
VRM_RENDER_MODE=3 timeout 900 ncu --set full --import-source on --clock-control none -k regex:render_lean_kernel -s 2 -c 1 -o gpurun_out/prof_r02b_lean_hash_orig -f \
    python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json > gpurun_out/r02b_ncu_lean_hash_orig.log 2>&1
ls -la gpurun_out/*r02b*
