#!/bin/bash
# refresh of the round-2 single-GPU artefacts after the last storage-solver changes
mkdir -p gpurun_out
P=gpurun_out/r2G
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -k "not long_horizon" 2>&1 | grep -v "^  \|^    " | tail -6 > ${P}_pytest.log
tail -3 ${P}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; head -c 330 ${P}_bench.json; echo; tail -2 ${P}_bench.err
timeout 300 python bench.py --steps 200 --warmup 30 --quick --no-cpu-baseline > ${P}_bench_long.json 2> ${P}_bench_long.err; head -c 330 ${P}_bench_long.json; echo
timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg3 --quick --no-cpu-baseline > ${P}_bench_cfg3.json 2> ${P}_bench_cfg3.err; head -c 330 ${P}_bench_cfg3.json; echo
timeout 300 python scripts/transient.py target 1 26 8,14 0.03 > ${P}_transient.log 2>&1; tail -3 ${P}_transient.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file ${P}_launches.csv python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > ${P}_ncu_bench.log 2>&1; wc -l ${P}_launches.csv
python scripts/prof_case.py 2000 3000 80000 20000 96 2 7 > ${P}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sto_warp|k_gen_predict|k_gemm" -s 29 -c 4 -f -o ${P}_top python scripts/prof_case.py 2000 3000 80000 20000 96 2 7 > ${P}_ncu_full.log 2>&1
tail -2 ${P}_ncu_full.log; ls -la ${P}_top.ncu-rep
