#!/bin/bash
# round-2 GPU call 2: new-boundary tests, full GPU suite, short-distance L2 prefetch probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2_pytest2_new.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest2_new.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py > gpurun_out/r2_pytest2_all.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest2_all.log
timeout 600 tools/ab.sh "base|X=1" "pf1_m01|SBLX_PF_MODE=1 SBLX_PF_MULT=0.1" "pf2_m01|SBLX_PF_MODE=2 SBLX_PF_MULT=0.1" "pf2_m005|SBLX_PF_MODE=2 SBLX_PF_MULT=0.05" "pf2_m02|SBLX_PF_MODE=2 SBLX_PF_MULT=0.2" > gpurun_out/r2_ab2.txt 2>&1
