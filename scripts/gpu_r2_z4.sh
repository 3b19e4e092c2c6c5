#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 20 --warmup 5 --quick > gpurun_out/r2z4_bench_weak.json 2> gpurun_out/r2z4_bench_weak.err; head -c 330 gpurun_out/r2z4_bench_weak.json; echo; tail -2 gpurun_out/r2z4_bench_weak.err
