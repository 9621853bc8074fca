#!/bin/bash
# round-2 GPU call 3: round-2 tests, bench.py in its new form (short), other configs (short), compute-sanitizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q > gpurun_out/r2_pytest3_new.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest3_new.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "retry_flat or contingenc or island" > gpurun_out/r2_pytest3_sel.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest3_sel.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_n1_short.json 2> gpurun_out/r2_bench_n1_short.err
echo "bench rc $?" >> gpurun_out/r2_bench_n1_short.err
timeout 600 python bench.py --config n1_2916_strong --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_n1_short.json 2> gpurun_out/r2_bench_c3_n1_short.err
timeout 900 python bench.py --config sweep_100k --scenarios 16384 --steps 1 --warmup 1 > gpurun_out/r2_bench_c4_16k.json 2> gpurun_out/r2_bench_c4_16k.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_short.json 2> gpurun_out/r2_bench_ref_short.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
