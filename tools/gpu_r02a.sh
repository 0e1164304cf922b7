#!/bin/bash
# Round 2, first GPU call: the whole GPU test tier (new build pipeline, full-size config parity), bench, L2 roofs, A/B of the
# two-phase nested traversal, launch lists and full captures of the kernels VERDICT r01 asked for.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
VRM_BUILD_TRACE=1 timeout 300 python tools/ncu_targets.py build3 vcs > gpurun_out/r02a_build_vcs.log 2>&1
VRM_BUILD_TRACE=1 timeout 300 python tools/ncu_targets.py build3 hashtable > gpurun_out/r02a_build_hash.log 2>&1
timeout 300 python tools/ncu_targets.py build3 vcs > gpurun_out/r02a_build_vcs_plain.log 2>&1
timeout 300 python tools/ncu_targets.py build3 hashtable > gpurun_out/r02a_build_hash_plain.log 2>&1
grep -h "build3" gpurun_out/r02a_build_*plain.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err || echo "bench failed"
tail -c 1500 gpurun_out/r02a_bench_n1.json; echo
timeout 300 python tools/l2_bw.py --out gpurun_out/r02a_l2_bw.json > gpurun_out/r02a_l2_bw.log 2>&1; cat gpurun_out/r02a_l2_bw.log
timeout 900 python tools/ab_variants.py main onephase > gpurun_out/r02a_ab.log 2>&1; cat gpurun_out/r02a_ab.log
cp gpurun_out/ab.json gpurun_out/r02a_ab.json
# launch lists (one pass each)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-baselines > gpurun_out/r02a_ncu_launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_build_vcs.csv \
    python tools/ncu_targets.py build3 vcs > gpurun_out/r02a_ncu_launches_build_vcs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_build_hash.csv \
    python tools/ncu_targets.py build3 hashtable > gpurun_out/r02a_ncu_launches_build_hash.log 2>&1
# full captures
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/prof_r02a_trace5_la -f \
    python tools/ncu_targets.py trace5 longestaxis > gpurun_out/r02a_ncu_trace5.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_r02a_orbit4_la -f \
    python tools/ncu_targets.py orbit4 longestaxis > gpurun_out/r02a_ncu_orbit4.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"radix_scatter|vcs_fill|radix_hist" -s 12 -c 3 -o gpurun_out/prof_r02a_build_vcs -f \
    python tools/ncu_targets.py build3 vcs > gpurun_out/r02a_ncu_build_vcs.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"cuckoo_insert" -s 1 -c 1 -o gpurun_out/prof_r02a_build_hash -f \
    python tools/ncu_targets.py build3 hashtable > gpurun_out/r02a_ncu_build_hash.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_r02a_hash_orig -f \
    python tools/explore.py --iters 2 --combos hashtable:original --out gpurun_out/x.json > gpurun_out/r02a_ncu_hash_orig.log 2>&1
ls -la gpurun_out/*r02a* | head -40
