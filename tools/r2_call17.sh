#!/bin/bash
# round-2 GPU call 17: resident-slot sweep (TLB reach / L2 locality against grid size) on the full 29 601-case N-1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 tools/ab_args.sh "slots_auto|--steps 1 --warmup 1" "slots_512|--steps 1 --warmup 1 --max-slots 512" "slots_1024|--steps 1 --warmup 1 --max-slots 1024" "slots_2048|--steps 1 --warmup 1 --max-slots 2048" "slots_3072|--steps 1 --warmup 1 --max-slots 3072" "slots_256|--steps 1 --warmup 1 --max-slots 256" > gpurun_out/r2_ab17.txt 2>&1
