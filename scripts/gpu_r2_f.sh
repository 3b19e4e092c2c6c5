#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_case.py 2000 3000 80000 20000 96 2 60 > gpurun_out/r2f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_sto_warp" -s 60 -c 1 -f -o gpurun_out/r2f_sto python scripts/prof_case.py 2000 3000 80000 20000 96 2 60 > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log; ls -la gpurun_out/r2f_sto.ncu-rep
