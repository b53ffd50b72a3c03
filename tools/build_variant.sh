#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=..]...  -> takzero_b200/build/variants/NAME.so (tuning aid)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p takzero_b200/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --cudart static -shared "$@" -o takzero_b200/build/variants/$name.so \
  takzero_b200/csrc/api.cu takzero_b200/csrc/kernels.cu takzero_b200/csrc/nn.cu takzero_b200/csrc/rnd.cu takzero_b200/csrc/comm.cu takzero_b200/csrc/model_file.cpp -ldl
echo built $name
