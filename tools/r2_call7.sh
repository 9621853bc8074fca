#!/bin/bash
# round-2 GPU call 7 (1 GPU): ncu evidence - launch list of the default bench command, per-launch metrics of one tick,
# --set full of k_front launches.  Numbers printed by runs under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --scenarios 2048 --steps 1 --warmup 0 --no-cpu-baseline --no-api-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-api-e2e > gpurun_out/r2_ncu_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum
timeout 1200 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_tick_metrics.csv \
  env SBLX_PROFILE_TICK=3 $B > gpurun_out/r2_ncu_tick.log 2>&1
timeout 1200 ncu --profile-from-start off --clock-control none --set full -k regex:k_front -c 14 -o /tmp/r2_k_front_set_full -f \
  env SBLX_PROFILE_TICK=3 $B > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/r2_k_front_set_full.ncu-rep --page raw --csv > gpurun_out/r2_k_front_set_full_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out
