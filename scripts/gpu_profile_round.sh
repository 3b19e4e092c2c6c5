#!/bin/bash
# round artefacts: plain bench, ncu launch list of the same command, full capture of the dominant kernel
set -o pipefail
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 30 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 720 -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 20 --warmup 30 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 400 gpurun_out/bench_short.json
python scripts/prof_case.py 2000 3000 80000 20000 96 2 60 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_sto_warp|k_gen_predict|k_gemm" -s 240 -c 4 -f -o gpurun_out/prof_r1_top \
    python scripts/prof_case.py 2000 3000 80000 20000 96 2 60 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
