#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_instantiations.py tests/test_gpu_parity.py -q -m gpu --timeout 900 -x 2>&1 | grep -v "^  \|^    " | tail -12 > gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
DOPF_LIB=$PWD/build/libdopf_stats.so timeout 600 python scripts/sto_stats.py target 0.03 1 1,3,8,16,40,120 > gpurun_out/r2g_stats.log 2>&1
grep "predict\|anchored" gpurun_out/r2g_stats.log | sed 's/anchors.*gt8_rounds/gt8/' | cut -c1-170
timeout 600 python scripts/transient.py target 1 40 3,8,16,40 0.03 > gpurun_out/r2g_transient.log 2>&1; tail -5 gpurun_out/r2g_transient.log
