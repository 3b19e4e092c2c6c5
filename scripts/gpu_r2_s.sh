#!/bin/bash
mkdir -p gpurun_out
for v in w4av w2 w1 dyn dyn1; do
  AB_PROF_AT=10,14 DOPF_LIB=$PWD/variants/libdopf_$v.so timeout 200 python scripts/ab2.py target k_sto_warp k_sto_fix 2>&1 | tail -3
done | tee gpurun_out/r2s_ab.log
