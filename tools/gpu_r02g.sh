#!/bin/bash
# Round 2, GPU call g: warp-uniform fast paths of the VCS + longest-axis render kernel (all-jump / all-null-region passes): GPU test tier,
# A/B against the build without them, capture of the new kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02g_pytest.log
tail -12 gpurun_out/r02g_pytest.log
timeout 1200 python tools/ab_variants.py prev@vcs:longestaxis nofast@vcs:longestaxis main@vcs:longestaxis nofast@vcs:longestaxis main@vcs:longestaxis > gpurun_out/r02g_ab.log 2>&1; cat gpurun_out/r02g_ab.log
cp gpurun_out/ab.json gpurun_out/r02g_ab.json
bash tools/gpu_capture.sh r02g_ncu_render_vcs_longestaxis render_kernel 3 terrain512_4k_vcs_longestaxis -- python bench.py --steps 2 --warmup 3 --no-baselines --single-view
du -sh gpurun_out
