#!/bin/bash
# round-2 GPU call 5: phase ablation of k_front (garbage results, timing only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/sparlectra.jl_b200/lib
A="SBLX_LIB=$L/libsblx_ablate.so"
M="SBLX_MODE_AFTER_BASE"
timeout 1500 tools/ab.sh "abl_base|$A" "abl_schurfma|$A $M=8" "abl_gather|$A $M=16" "abl_cstore|$A $M=32" \
  "abl_schur_all|$A $M=56" "abl_ingestrest|$A $M=64" "abl_subst|$A $M=128" "abl_ustore|$A $M=256" "abl_chain|$A $M=512" \
  "abl_all_but_schur|$A $M=960" "abl_everything|$A $M=1016" "abl_gather_store|$A $M=48" > gpurun_out/r2_ab5.txt 2>&1
