#!/bin/bash
# round-2 GPU call 1: GPU tests of the default build + A/B of the L2-prefetch and cp.async ingest experiments
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest1.log
L=$PWD/sparlectra.jl_b200/lib
timeout 900 tools/ab.sh "base|X=1" "pf1|SBLX_PF_MODE=1" "pf2|SBLX_PF_MODE=2" "pf4|SBLX_PF_MODE=4" "pf5|SBLX_PF_MODE=5" \
  "pf2_m05|SBLX_PF_MODE=2 SBLX_PF_MULT=0.5" "pf2_m2|SBLX_PF_MODE=2 SBLX_PF_MULT=2" "pf1_m2|SBLX_PF_MODE=1 SBLX_PF_MULT=2" \
  "cpasync|SBLX_LIB=$L/libsblx_cpasync.so" "cpasync_pf2|SBLX_LIB=$L/libsblx_cpasync.so SBLX_PF_MODE=2" "cpasync_pf5|SBLX_LIB=$L/libsblx_cpasync.so SBLX_PF_MODE=5" \
  "base2|X=1" > gpurun_out/r2_ab1.txt 2>&1
SBLX_LIB=$L/libsblx_cpasync.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size" > gpurun_out/r2_pytest_cpasync.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest_cpasync.log
