#!/bin/bash
mkdir -p gpurun_out
AB_PROF_AT=8,14 timeout 200 python scripts/ab2.py target k_verify k_gen_fix k_sto_collect k_sto_fix 2>&1 | tail -3 | tee gpurun_out/r2w_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_partition.py -q -m gpu --timeout 600 -x 2>&1 | tail -3
