#!/bin/bash
# round-2 GPU call 15: did the bulk L2 prefetch actually land?  L2 hit rate / DRAM bytes of k_front with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --scenarios 2048 --steps 1 --warmup 0 --no-cpu-baseline --no-api-e2e"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for v in "off:0:1" "pf2_m01:2:0.1" "pf2_m1:2:1" "pf1_m01:1:0.1"; do
  name=${v%%:*}; rest=${v#*:}; mode=${rest%%:*}; mult=${rest#*:}
  timeout 900 ncu --profile-from-start off --clock-control none --metrics $M -k regex:k_front --csv --log-file gpurun_out/r2_pfcheck_$name.csv \
    env SBLX_PROFILE_TICK=3 SBLX_PF_MODE=$mode SBLX_PF_MULT=$mult $B > gpurun_out/r2_pfcheck_$name.log 2>&1
done
