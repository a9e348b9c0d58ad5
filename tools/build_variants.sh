#!/bin/bash
# tools/build_variants.sh NAME "EXTRA nvcc flags" [NAME "flags" ...] -- builds build/variants/NAME.so for tools/kbench.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/build/variants"
cd "$ROOT/lzma_rs_b200/csrc"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $flags -diag-suppress 186 -O3 -std=c++17 -lineinfo \
      -Xcompiler -fPIC -I../../include -I. -shared -o "$ROOT/build/variants/$name.so" \
      lzb_kernels.cu lzb_encode_kernels.cu lzb_host.cu lzb_plan.cpp && echo "built $name" ) &
done
wait
