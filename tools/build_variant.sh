#!/bin/bash
# Builds a compile-time variant of liborbx.so into variants/liborbx_<name>.so (git-ignored, travels to the GPU box).
# usage: bash tools/build_variant.sh <name> [-DFLAG=..]...
NAME=$1; shift
mkdir -p variants
cd openmavis_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function,-ffp-contract=off "$@" -shared -o ../../variants/liborbx_$NAME.so orbx_extract.cu orbx_match.cu -lcudart
