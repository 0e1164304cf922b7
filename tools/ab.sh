# A/B on a GPU box: GPU parity tests, then the four combinations on the 512^3 terrain at 4K with both render modes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
for mode in 1 2; do
VRM_RENDER_MODE=$mode python tools/explore.py --iters 7 --out gpurun_out/explore_mode$mode.json 2>&1 | grep -E '^\{' | MODE=$mode python -c "
import sys, json, os
for l in sys.stdin:
    r=json.loads(l); print('MODE', os.environ['MODE'], '(1=nested,2=flatloop)', r['storage'], r['algo'], round(r['ms'],3), 'ms', round(r['mrays']), 'Mrays/s')
"
done
