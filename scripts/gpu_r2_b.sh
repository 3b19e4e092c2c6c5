#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
  echo "== debug_flags $dbg"; timeout 300 python scripts/diag_storage.py 40 60 80 10 130 10 12 3 0.3 $dbg 2>&1 | tail -22
done > gpurun_out/r2b_diag.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_instantiations.py -q -m gpu --timeout 900 2>&1 | grep -v "^  \|^E   \|^    " | tail -40 > gpurun_out/r2b_pytest_new.log
timeout 900 python scripts/param_scan.py cfg3 800 0.3:10 0.3:1 0.1:3.33 0.03:1 0.03:0.1 1:3.3 > gpurun_out/r2b_scan_cfg3.log 2>&1
timeout 900 python scripts/param_scan.py cfg2 4000 0.3:10 0.3:1 0.1:3.33 0.03:1 1:33 1:3.3 > gpurun_out/r2b_scan_cfg2.log 2>&1
tail -30 gpurun_out/r2b_diag.log; tail -15 gpurun_out/r2b_pytest_new.log
