#!/bin/bash
# Tuning builds of libshipenv.so: tools/build_variant.sh NAME [extra nvcc flags...] -> gpurun_variants/libshipenv_NAME.so
# (run with SHIPENV_LIB=gpurun_variants/libshipenv_NAME.so; the directory is git-ignored and travels with gpurun)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared \
  "$@" -o gpurun_variants/libshipenv_$name.so shippingenv_b200/csrc/shipenv_abi.cu
echo built gpurun_variants/libshipenv_$name.so
