#!/bin/bash
mkdir -p gpurun_out
AB_PROF_AT=8 timeout 200 python scripts/ab2.py target k_sto_warp k_sto_fix 2>&1 | tail -2 | tee gpurun_out/r2w2_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_instantiations.py -q -m gpu --timeout 600 -x -k "not long_horizon and not J8 and not J6" 2>&1 | tail -3
