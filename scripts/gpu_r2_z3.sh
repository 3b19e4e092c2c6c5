#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2z3_bench_weak.json 2> gpurun_out/r2z3_bench_weak.err; head -c 330 gpurun_out/r2z3_bench_weak.json; echo
timeout 300 $TR --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --quick > gpurun_out/r2z3_bench_strong.json 2> gpurun_out/r2z3_bench_strong.err; head -c 330 gpurun_out/r2z3_bench_strong.json; echo
