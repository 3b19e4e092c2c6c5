#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_long_horizon.py tests/test_gpu_ptdf.py tests/test_gpu_batch.py tests/test_gpu_partition.py tests/test_gpu_mirror.py -q -m gpu --timeout 900 2>&1 | grep -v "^  \|^    " | tail -25 > gpurun_out/r2n_pytest.log
tail -8 gpurun_out/r2n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python scripts/converge_probe.py 30 45 60 12 12 400000 0.3:10 0.03:1 0.1:1 > gpurun_out/r2n_converge_small.log 2>&1; cat gpurun_out/r2n_converge_small.log | grep -v "it [0-9]*0000 res" | tail -12
timeout 600 python scripts/transient.py target 1 26 6,10,14 0.03 > gpurun_out/r2n_transient.log 2>&1; tail -4 gpurun_out/r2n_transient.log
