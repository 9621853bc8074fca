#!/bin/bash
# round-2 GPU call 6 (2 GPUs): round-2 tests incl. the in-library multi-GPU path, full suite, bounds-checked build, bench at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/sparlectra.jl_b200/lib
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_gpus6.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q > gpurun_out/r2_pytest6_new.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest6_new.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py > gpurun_out/r2_pytest6_all.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest6_all.log
SBLX_LIB=$L/libsblx_check.so timeout 600 python tools/sanitize_run.py > gpurun_out/r2_bounds_check.txt 2>&1
echo "rc $?" >> gpurun_out/r2_bounds_check.txt
SBLX_LIB=$L/libsblx_check.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not full_size and not 54x54 and not concurrent" >> gpurun_out/r2_bounds_check.txt 2>&1
echo "pytest(check build) rc $?" >> gpurun_out/r2_bounds_check.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2_short.json 2> gpurun_out/r2_bench_n2_short.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --config n1n2_10k --scenarios 8192 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2_weak_short.json 2> gpurun_out/r2_bench_n2_weak_short.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --config n1_2916_strong --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_n2_short.json 2> gpurun_out/r2_bench_c3_n2_short.err
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --config sweep_100k --scenarios 16384 --steps 1 --warmup 1 > gpurun_out/r2_bench_c4_n2_short.json 2> gpurun_out/r2_bench_c4_n2_short.err
