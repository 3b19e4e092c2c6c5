#!/bin/bash
mkdir -p gpurun_out
DOPF_LIB=$PWD/variants/libdopf_stats.so timeout 300 python scripts/sto_stats.py target 0.03 1 6,8,10,14,20,24 2>&1 | grep -E "STRAG|predict" | cut -c1-220 > gpurun_out/r2t_strag.log; head -120 gpurun_out/r2t_strag.log
