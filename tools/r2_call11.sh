#!/bin/bash
# round-2 GPU call 11 (1 GPU): final verification of the tree - full GPU suite, smoke, stress tools, latencies, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/sparlectra.jl_b200/lib
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest11.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke11.log 2>&1
echo "smoke rc $?" >> gpurun_out/r2_smoke11.log
timeout 300 python tools/latency_small.py > gpurun_out/r2_latency_small.txt 2>&1
SBLX_LIB=$L/libsblx_check.so timeout 600 python tools/sanitize_run.py > gpurun_out/r2_bounds_check11.txt 2>&1
echo "rc $?" >> gpurun_out/r2_bounds_check11.txt
(timeout 600 python tools/stress.py 16 211; timeout 600 python tools/stress_islands.py 24 212; timeout 600 python tools/stress_keywords.py 16 213; timeout 600 python tools/stress_injections.py 12 214) > gpurun_out/r2_stress11.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
echo "bench rc $?" >> gpurun_out/r2_bench_n1_final.err
