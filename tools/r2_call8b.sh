#!/bin/bash
# round-2 GPU call 8b (8 GPUs): in-library multi-GPU tests and the bench configs at N = 8 (and the default at N = 4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_gpus8.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "two_gpus or two_processes or sharded or shard_bounds" > gpurun_out/r2_pytest8b_multi.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest8b_multi.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
timeout 600 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --config n1n2_10k --steps 3 --warmup 2 > gpurun_out/r2_bench_n8_weak.json 2> gpurun_out/r2_bench_n8_weak.err
timeout 600 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --config n1_2916_strong --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_n8.json 2> gpurun_out/r2_bench_c3_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --config sweep_100k --steps 2 --warmup 1 > gpurun_out/r2_bench_c4_n8.json 2> gpurun_out/r2_bench_c4_n8.err
