#!/bin/bash
# A/B builds of the same ABI with other compile-time constants: tools/build_alt.sh NAME -DORZ_V7_LEAD=64 ...  ->  build_alt/liborz_NAME.so
# (git-ignored; travels to the GPU box; select with ORZ_B200_LIB=build_alt/liborz_NAME.so)
set -e
name=$1; shift
cd "$(dirname "$0")/../orz_b200/csrc"
mkdir -p ../../build_alt
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o ../../build_alt/liborz_$name.so capi.cu encode_kernels.cu decode_kernels.cu
