#!/bin/bash
# Builds the reference-shaped multithreaded C++ CPU baseline (oracle/cpu_ref/cpu_ref.cpp) -> oracle/_build/libcpuref.so.
# Test / measurement infrastructure: only tests/, bench.py's cpu_baseline leg and `bench.py --impl reference` load it.
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p _build
g++ -O3 -march=native -std=c++17 -fPIC -shared -pthread cpu_ref/cpu_ref.cpp -o _build/libcpuref.so
echo "built $(pwd)/_build/libcpuref.so"
