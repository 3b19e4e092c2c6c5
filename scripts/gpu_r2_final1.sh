#!/bin/bash
# round-2 artefacts on one GPU: full GPU test suite, bench lines, ncu launch list + full capture of the top kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 --durations=12 2>&1 | grep -v "^  \|^    " | tail -30 > gpurun_out/r2f_pytest.log
tail -22 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 300 gpurun_out/r2f_bench.json; tail -2 gpurun_out/r2f_bench.err
timeout 300 python bench.py --steps 200 --warmup 30 --quick --no-cpu-baseline > gpurun_out/r2f_bench_long.json 2> gpurun_out/r2f_bench_long.err; head -c 330 gpurun_out/r2f_bench_long.json; echo
for wl in cfg2 cfg3 cfg4; do timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --quick --no-cpu-baseline > gpurun_out/r2f_bench_$wl.json 2> gpurun_out/r2f_bench_$wl.err; head -c 330 gpurun_out/r2f_bench_$wl.json; echo; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1; wc -l gpurun_out/r2f_launches.csv
python scripts/prof_case.py 2000 3000 80000 20000 96 2 7 > gpurun_out/r2f_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sto_warp|k_gen_predict|k_gemm" -s 29 -c 4 -f -o gpurun_out/r2f_top python scripts/prof_case.py 2000 3000 80000 20000 96 2 7 > gpurun_out/r2f_ncu_full.log 2>&1
tail -2 gpurun_out/r2f_ncu_full.log; ls -la gpurun_out/r2f_top.ncu-rep
