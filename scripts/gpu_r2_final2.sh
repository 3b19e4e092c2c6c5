#!/bin/bash
# round-2 final artefacts on one GPU: full GPU test suite, bench lines, transient log, ncu launch list + full capture of the top kernels
mkdir -p gpurun_out
P=gpurun_out/r2F
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 --durations=8 2>&1 | grep -v "^  \|^    " | tail -24 > ${P}_pytest.log
tail -14 ${P}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; head -c 330 ${P}_bench.json; echo; tail -2 ${P}_bench.err
timeout 300 python bench.py --steps 200 --warmup 30 --quick --no-cpu-baseline > ${P}_bench_long.json 2> ${P}_bench_long.err; head -c 330 ${P}_bench_long.json; echo
for wl in cfg2 cfg3 cfg4; do timeout 300 python bench.py --steps 20 --warmup 5 --workload $wl --quick --no-cpu-baseline > ${P}_bench_$wl.json 2> ${P}_bench_$wl.err; head -c 330 ${P}_bench_$wl.json; echo; done
timeout 300 python bench.py --steps 20 --warmup 5 --path partitioned --quick --no-cpu-baseline > ${P}_bench_part1.json 2> ${P}_bench_part1.err; head -c 330 ${P}_bench_part1.json; echo
timeout 300 python bench.py --steps 20 --warmup 5 --path partitioned --comm library --quick --no-cpu-baseline > ${P}_bench_part1_lib.json 2> ${P}_bench_part1_lib.err; head -c 330 ${P}_bench_part1_lib.json; echo
timeout 300 python scripts/transient.py target 1 26 8,14 0.03 > ${P}_transient.log 2>&1; tail -3 ${P}_transient.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file ${P}_launches.csv python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > ${P}_ncu_bench.log 2>&1; wc -l ${P}_launches.csv
python scripts/prof_case.py 2000 3000 80000 20000 96 2 7 > ${P}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sto_warp|k_gen_predict|k_gemm" -s 29 -c 4 -f -o ${P}_top python scripts/prof_case.py 2000 3000 80000 20000 96 2 7 > ${P}_ncu_full.log 2>&1
tail -2 ${P}_ncu_full.log; ls -la ${P}_top.ncu-rep
