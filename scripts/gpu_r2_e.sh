#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | grep -v "^  \|^    " | tail -30 > gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
DOPF_LIB=$PWD/build/libdopf_stats.so timeout 600 python scripts/sto_stats.py target 0.03 1 1,2,3,5,8,12,16,25,80 > gpurun_out/r2e_stats.log 2>&1
grep predict gpurun_out/r2e_stats.log | cut -c1-60,330-
timeout 600 python scripts/transient.py target 1 40 3,8,16,40 0.03 > gpurun_out/r2e_transient.log 2>&1; tail -5 gpurun_out/r2e_transient.log
timeout 600 python scripts/transient.py target 1 40 "" 0.3 > gpurun_out/r2e_transient_g03.log 2>&1; tail -1 gpurun_out/r2e_transient_g03.log
