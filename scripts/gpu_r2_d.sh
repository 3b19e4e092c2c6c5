#!/bin/bash
mkdir -p gpurun_out
DOPF_LIB=$PWD/build/libdopf_stats.so timeout 600 python scripts/sto_stats.py target 0.03 1 1,2,3,5,8,12,16,25,40,80,200 > gpurun_out/r2d_stats.log 2>&1
tail -40 gpurun_out/r2d_stats.log
timeout 1800 python -m pytest tests/test_gpu_instantiations.py tests/test_gpu_mirror.py -q -m gpu --timeout 900 2>&1 | grep -v "^  \|^E   \|^    " | tail -30 > gpurun_out/r2d_pytest.log
tail -12 gpurun_out/r2d_pytest.log
timeout 600 python scripts/transient.py target 1 30 "" > gpurun_out/r2d_transient.log 2>&1; tail -3 gpurun_out/r2d_transient.log
