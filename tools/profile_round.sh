#!/bin/bash
# Round-end measurement set on a GPU box (1 GPU):  bash tools/profile_round.sh <tag>
#   bench line (native + reference arm), the ncu launch list of the bench command, and one `--set full` capture of the
#   dominant kernel of the headline combination and of the hash-table kernel.  Nothing printed under ncu is a bench value.
tag=${1:-rXX}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err || echo "bench failed"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err || echo "reference arm failed"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 5 --warmup 3 --no-baselines > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_${tag}_vcs_la -f \
    python tools/explore.py --iters 2 --combos vcs:longestaxis --out gpurun_out/x.json > gpurun_out/ncu_${tag}_vcs_la.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_${tag}_hash_orig -f \
    python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json > gpurun_out/ncu_${tag}_hash_orig.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_${tag}_vcs_orig -f \
    python tools/explore.py --iters 2 --combos vcs:original --out gpurun_out/x.json > gpurun_out/ncu_${tag}_vcs_orig.log 2>&1
tail -c 600 gpurun_out/bench_${tag}_n1.json; echo; cat gpurun_out/bench_${tag}_reference_arm.json
