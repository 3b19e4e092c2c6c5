#!/bin/bash
mkdir -p gpurun_out
DOPF_DEBUG_FLAGS=32 timeout 300 python scripts/transient.py target 1 60 26,60 0.03 > gpurun_out/r2o_seq.log 2>&1; tail -3 gpurun_out/r2o_seq.log | cut -c1-700
