#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py tests/test_gpu_partition.py -q -m gpu --timeout 600 2>&1 | grep -v "^    " | tail -30 > gpurun_out/r2j_tests.log
tail -6 gpurun_out/r2j_tests.log
timeout 600 python scripts/batch_probe.py 128 20 > gpurun_out/r2j_probe128.log 2>&1; cat gpurun_out/r2j_probe128.log | cut -c1-900
timeout 600 python scripts/batch_probe.py 1024 10 > gpurun_out/r2j_probe1024.log 2>&1; tail -3 gpurun_out/r2j_probe1024.log | cut -c1-900
DOPF_DEBUG_FLAGS=8 timeout 600 python scripts/transient.py target 1 40 40 0.03 > gpurun_out/r2j_flat.log 2>&1; tail -2 gpurun_out/r2j_flat.log
DOPF_DEBUG_FLAGS=16 timeout 600 python scripts/transient.py target 1 40 40 0.03 > gpurun_out/r2j_node.log 2>&1; tail -2 gpurun_out/r2j_node.log
timeout 600 python scripts/transient.py cfg2 1 60 50 0.03 > gpurun_out/r2j_cfg2.log 2>&1; tail -2 gpurun_out/r2j_cfg2.log
