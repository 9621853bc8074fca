#!/bin/bash
# round-2 GPU call 4: phase ablation of k_front (garbage results, timing only) and occupancy variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/sparlectra.jl_b200/lib
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "names_from_device" > gpurun_out/r2_pytest4.log 2>&1
A="SBLX_LIB=$L/libsblx_ablate.so"
timeout 1500 tools/ab.sh "base|X=1" "abl_base|$A" "abl_schurfma|$A SBLX_PF_MODE=8" "abl_gather|$A SBLX_PF_MODE=16" "abl_cstore|$A SBLX_PF_MODE=32" \
  "abl_schur_all|$A SBLX_PF_MODE=56" "abl_ingestrest|$A SBLX_PF_MODE=64" "abl_subst|$A SBLX_PF_MODE=128" "abl_ustore|$A SBLX_PF_MODE=256" "abl_chain|$A SBLX_PF_MODE=512" \
  "abl_all_but_schur|$A SBLX_PF_MODE=960" "abl_everything|$A SBLX_PF_MODE=1016" \
  "occ32g24|SBLX_LIB=$L/libsblx_occ32g24.so" "occ64_14|SBLX_LIB=$L/libsblx_occ64_14.so" "occ256_3|SBLX_LIB=$L/libsblx_occ256_3.so" "base2|X=1" > gpurun_out/r2_ab4.txt 2>&1
