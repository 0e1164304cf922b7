#!/bin/bash
# Round 2, fourth GPU call: forms of the shadow kernel (nested / state machine / lane-level refill), parity of each through the GPU tests.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
for f in 2; do
  ( time VRM_SHADOW_FORM=$f timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_configs_gpu.py -m gpu -q -x ) > gpurun_out/r02e_pytest_form$f.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02e_pytest_form$f.log
  tail -6 gpurun_out/r02e_pytest_form$f.log
done
timeout 1500 python tools/ab_variants.py prev main main:shadow=0 main:shadow=1 main:shadow=2 rf4:shadow=2 rf16:shadow=2 rf24:shadow=2 > gpurun_out/r02e_ab.log 2>&1; cat gpurun_out/r02e_ab.log
cp gpurun_out/ab.json gpurun_out/r02e_ab.json
for spec in prev:x main:0 main:1 main:2; do
  lib=${spec%%:*}; f=${spec#*:}
  if [ $lib = prev ]; then export VRM_B200_LIB=$PWD/voxelraymarcher_b200/variants/libvrm_prev.so; else unset VRM_B200_LIB; fi
  if [ $f = x ]; then unset VRM_SHADOW_FORM; else export VRM_SHADOW_FORM=$f; fi
  for a in longestaxis original; do
    timeout 300 python tools/ncu_targets.py trace5 $a 2>&1 | grep trace5 | sed "s/^/$spec: /"
    timeout 300 python tools/ncu_targets.py orbit4 $a 2>&1 | grep "orbit4" | sed "s/^/$spec: /"
  done
done > gpurun_out/r02e_cfg45.log 2>&1
unset VRM_B200_LIB VRM_SHADOW_FORM
cat gpurun_out/r02e_cfg45.log
export VRM_SHADOW_FORM=2
bash tools/gpu_capture.sh r02e_ncu_shadow_refill_vcs_longestaxis shadow_refill_kernel 3 - -- python bench.py --steps 2 --warmup 3 --no-baselines --single-view
bash tools/gpu_capture.sh r02e_ncu_shadow_refill_hashtable_original shadow_refill_kernel 3 - -- python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json
bash tools/gpu_capture.sh r02e_ncu_render_hashtable_original render_kernel 3 - -- python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json
rm -f gpurun_out/x.json
du -sh gpurun_out
