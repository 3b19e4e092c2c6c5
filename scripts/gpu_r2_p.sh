#!/bin/bash
mkdir -p gpurun_out
DOPF_LIB=$PWD/build/libdopf_stats.so timeout 300 python scripts/sto_stats.py target 0.03 1 2,4,6,8,10,12,14,16,20,24 2>&1 | grep predict | sed 's/anchors.*gt8_rounds/gt8/' | cut -c1-200 > gpurun_out/r2p_stats.log; cat gpurun_out/r2p_stats.log
timeout 300 python scripts/transient.py target 1 26 8,14 0.03 > gpurun_out/r2p_new.log 2>&1; tail -3 gpurun_out/r2p_new.log | cut -c1-400
DOPF_DEBUG_FLAGS=64 timeout 300 python scripts/transient.py target 1 26 8,14 0.03 > gpurun_out/r2p_old.log 2>&1; tail -3 gpurun_out/r2p_old.log | cut -c1-400
