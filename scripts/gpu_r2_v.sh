#!/bin/bash
mkdir -p gpurun_out
DOPF_LIB=$PWD/variants/libdopf_stats.so timeout 300 python scripts/strag_time.py 30 2>&1 | grep STIME > gpurun_out/r2v_stime.log; wc -l gpurun_out/r2v_stime.log
