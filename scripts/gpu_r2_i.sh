#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -q -m gpu --timeout 600 2>&1 | grep -v "^    " | tail -30 > gpurun_out/r2i_batch.log
tail -12 gpurun_out/r2i_batch.log
timeout 600 python scripts/batch_probe.py 128 20 > gpurun_out/r2i_probe128.log 2>&1; cat gpurun_out/r2i_probe128.log | cut -c1-900
timeout 600 python scripts/batch_probe.py 1024 10 > gpurun_out/r2i_probe1024.log 2>&1; cat gpurun_out/r2i_probe1024.log | cut -c1-900
