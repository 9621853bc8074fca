#!/bin/bash
# round-2 GPU call 13 (1 GPU): look-ahead ticks for small batches - full suite, latencies with and without, bounds-checked build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/sparlectra.jl_b200/lib
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest13.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_pytest13.log
(echo "# look-ahead ticks (default)"; timeout 300 python tools/latency_small.py; echo "# SBLX_NO_LOOKAHEAD=1 (synchronous ticks)"; SBLX_NO_LOOKAHEAD=1 timeout 300 python tools/latency_small.py) > gpurun_out/r2_latency_small.txt 2>&1
SBLX_LIB=$L/libsblx_check.so timeout 600 python tools/sanitize_run.py > gpurun_out/r2_bounds_check13.txt 2>&1
echo "rc $?" >> gpurun_out/r2_bounds_check13.txt
SBLX_LIB=$L/libsblx_check.so timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -k "not full_size and not 54x54 and not concurrent and not 10k and not two_gpus and not two_processes" >> gpurun_out/r2_bounds_check13.txt 2>&1
echo "pytest(check build) rc $?" >> gpurun_out/r2_bounds_check13.txt
(timeout 600 python tools/stress.py 12 311; timeout 600 python tools/stress_islands.py 16 312; timeout 600 python tools/stress_keywords.py 12 313) > gpurun_out/r2_stress13.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke13.log 2>&1
