#!/bin/bash
# build the working tree's library with extra nvcc flags into variants/libdopf_<name>.so (A/B timing with DOPF_LIB=...)
set -e
name=$1; shift
mkdir -p variants
S=decentralopf.jl_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --shared -cudart static "$@" \
    -o "variants/libdopf_$name.so" $S/dopf_kernels.cu $S/dopf_api.cu $S/dopf_ptdf.cu -ldl
echo "variants/libdopf_$name.so"
