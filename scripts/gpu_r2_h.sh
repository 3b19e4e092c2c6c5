#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py -q -m gpu --timeout 600 2>&1 | grep -v "^    " | tail -40 > gpurun_out/r2h_batch.log
tail -25 gpurun_out/r2h_batch.log
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x --deselect tests/test_gpu_batch.py 2>&1 | grep -v "^  \|^    " | tail -12 > gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
