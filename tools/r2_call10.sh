#!/bin/bash
# round-2 GPU call 10: early-gather Schur variants (A/B) + parity of the variant builds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/sparlectra.jl_b200/lib
timeout 900 tools/ab.sh "base|X=1" "early|SBLX_LIB=$L/libsblx_early.so" "early_small|SBLX_LIB=$L/libsblx_early_small.so" "early_both|SBLX_LIB=$L/libsblx_early_both.so" "base2|X=1" > gpurun_out/r2_ab10.txt 2>&1
SBLX_LIB=$L/libsblx_early.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not concurrent" > gpurun_out/r2_pytest10_early.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest10_early.log
