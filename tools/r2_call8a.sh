#!/bin/bash
# round-2 GPU call 8a (1 GPU): round-2 tests with the refinement path, bounds-checked build, nose stress, full N=1 bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/sparlectra.jl_b200/lib
timeout 1200 python -m pytest tests/test_gpu_round2.py -m gpu -q > gpurun_out/r2_pytest8_new.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest8_new.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py > gpurun_out/r2_pytest8_all.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest8_all.log
SBLX_LIB=$L/libsblx_check.so timeout 600 python tools/sanitize_run.py > gpurun_out/r2_bounds_check.txt 2>&1
echo "rc $?" >> gpurun_out/r2_bounds_check.txt
SBLX_LIB=$L/libsblx_check.so timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -k "not full_size and not 54x54 and not concurrent and not 10k and not two_gpus and not two_processes" >> gpurun_out/r2_bounds_check.txt 2>&1
echo "pytest(check build) rc $?" >> gpurun_out/r2_bounds_check.txt
timeout 600 python tools/stress_nose.py 54 48 > gpurun_out/r2_stress_nose.txt 2>&1
timeout 900 python tools/stress_nose.py 100 32 >> gpurun_out/r2_stress_nose.txt 2>&1
timeout 300 python tools/latency_small.py > gpurun_out/r2_latency_small.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc $?" >> gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --config n1_2916_strong --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
timeout 900 python bench.py --config sweep_100k --steps 1 --warmup 1 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
