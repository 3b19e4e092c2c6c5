#!/bin/bash
# round 2, call A: new parity tests + cold-start transient data
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
timeout 1500 python -m pytest tests/test_gpu_instantiations.py -q -m gpu -x --timeout 900 2>&1 | tail -25 | tee gpurun_out/r2a_pytest_new.log
timeout 600 python scripts/transient.py target 1 60 3,8,16,40 > gpurun_out/r2a_transient_w1.log 2>&1
timeout 600 python scripts/transient.py target 10 60 3,8,16,40 > gpurun_out/r2a_transient_w10.log 2>&1
tail -8 gpurun_out/r2a_transient_w1.log; tail -8 gpurun_out/r2a_transient_w10.log
