#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_instantiations.py tests/test_gpu_batch.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -x 2>&1 | grep -v "^  \|^    " | tail -12 > gpurun_out/r2u_pytest.log
tail -5 gpurun_out/r2u_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --path partitioned --quick --no-cpu-baseline > gpurun_out/r2u_bench_part1.json 2> gpurun_out/r2u_bench_part1.err; head -c 700 gpurun_out/r2u_bench_part1.json; tail -3 gpurun_out/r2u_bench_part1.err
