#!/bin/bash
# N = 8: end-to-end forms for page-locked frames (SM stores vs copy engine bands) and per-warp stores for peer frames
cd "$(dirname "$0")/.."
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-baselines > gpurun_out/n8_e2e_$tag.json 2> gpurun_out/n8_e2e_$tag.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/n8_e2e_$tag.json"))
    print("$tag", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "pageable", round(d["e2e"]["pageable_caller_buffer"]["value"], 1), "verified", d.get("exchange_verified"))
except Exception as e:
    print("$tag failed", e)
PY
}
run default A=1
run dma4 VRM_PINNED_DMA=4
run dma2 VRM_PINNED_DMA=2
run wsremote VRM_WSTORE_REMOTE=1
