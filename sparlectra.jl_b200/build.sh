#!/bin/bash
# Builds libsblx.so (hand-written sm_100a kernels + C ABI) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")/csrc"
mkdir -p ../lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 -shared \
  engine.cu symbolic.cpp matpower.cpp -o ../lib/libsblx.so "$@"
echo "built $(cd ../lib && pwd)/libsblx.so"
