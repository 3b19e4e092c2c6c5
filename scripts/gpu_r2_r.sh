#!/bin/bash
mkdir -p gpurun_out
for v in head w4av w12av w12 w8av w6av; do
  DOPF_LIB=$PWD/variants/libdopf_$v.so timeout 200 python scripts/ab2.py target k_sto_warp k_sto_fix 2>&1 | tail -1
done | tee gpurun_out/r2r_ab.log
