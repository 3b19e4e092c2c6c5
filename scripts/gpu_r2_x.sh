#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_slack_rows|k_verify|k_inject|k_compact" -s 35 -c 5 -f -o gpurun_out/r2x_corr python scripts/prof_case.py 2000 3000 80000 20000 96 2 7 > gpurun_out/r2x_ncu.log 2>&1
tail -3 gpurun_out/r2x_ncu.log; ls -la gpurun_out/r2x_corr.ncu-rep
