mkdir -p gpurun_out
for v in old main; do
  if [ $v = old ]; then export VRM_B200_LIB=$PWD/voxelraymarcher_b200/variants/libvrm_old.so; else unset VRM_B200_LIB; fi
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__warps_active.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:render_kernel -c 3 --csv --log-file gpurun_out/ncu_cmp_$v.csv python tools/explore.py --iters 1 --combos vcs:original --out gpurun_out/x.json > gpurun_out/ncu_cmp_$v.log 2>&1
done
