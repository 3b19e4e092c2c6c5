#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | grep -v "^  \|^    " | tail -12 > gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 2500 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --path partitioned --quick --no-cpu-baseline > gpurun_out/r2k_bench_part1.json 2> gpurun_out/r2k_bench_part1.err; head -c 600 gpurun_out/r2k_bench_part1.json; tail -3 gpurun_out/r2k_bench_part1.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg4 --quick > gpurun_out/r2k_bench_cfg4.json 2> gpurun_out/r2k_bench_cfg4.err; head -c 500 gpurun_out/r2k_bench_cfg4.json; tail -3 gpurun_out/r2k_bench_cfg4.err
