#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/transient.py target 1 26 8,14 0.03 > gpurun_out/r2q_new.log 2>&1; tail -3 gpurun_out/r2q_new.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -q -m gpu --timeout 600 -x 2>&1 | tail -3
