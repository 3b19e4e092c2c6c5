#!/bin/bash
mkdir -p gpurun_out
# storage 1, iteration 5 of the failing J6 case: dump the rounds of the hinge solver
DBG=$(( 4 + (1<<8) + (5<<20) ))
DOPF_LIB=$PWD/build/libdopf_dbg.so timeout 300 python scripts/diag_storage.py 40 60 80 10 130 10 5 3 0.3 $DBG > gpurun_out/r2c_dump.log 2>&1
tail -5 gpurun_out/r2c_dump.log
timeout 600 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_partition.py -q -m gpu --timeout 600 2>&1 | grep -v "^  \|^    " | tail -40 > gpurun_out/r2c_pytest.log
tail -12 gpurun_out/r2c_pytest.log
timeout 300 python -m pytest "tests/test_gpu_instantiations.py::test_benchmarked_grid_T96" "tests/test_gpu_instantiations.py::test_gemm_64_row_tiles_and_split_k" -q -m gpu 2>&1 | grep "DopfError\|rc=" | head -5
