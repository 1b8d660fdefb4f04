#!/bin/bash
# Builds schwarzwald_b200/libswgpu.so for sm_100a (nvcc cross-compiles without a GPU).
# -fmad=false: the reference is built without FMA contraction; every a*b+c must round twice.
set -e
cd "$(dirname "$0")/schwarzwald_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC -shared ${SWGPU_NVCC_EXTRA} \
  -o ${SWGPU_OUT:-../libswgpu.so} kernels_index_sort.cu kernels_sampling.cu kernels_shard.cu kernels_payload.cu kernels_store.cu tiler.cu multi.cu
echo "built schwarzwald_b200/libswgpu.so"
