#!/bin/bash
# build the library of a git revision into variants/libdopf_<name>.so (A/B timing on one box with DOPF_LIB=...)
set -e
ref=$1; name=$2; shift 2
tmp=$(mktemp -d)
git archive "$ref" decentralopf.jl_b200/csrc include | tar -x -C "$tmp"
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --shared -cudart static "$@" \
    -o "variants/libdopf_$name.so" "$tmp/decentralopf.jl_b200/csrc/dopf_kernels.cu" "$tmp/decentralopf.jl_b200/csrc/dopf_api.cu"
rm -rf "$tmp"
echo "variants/libdopf_$name.so"
