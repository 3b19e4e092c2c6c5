#!/bin/bash
# 2-GPU call: partitioned correctness (torch-owned and library-owned NCCL) + weak / strong scaling bench lines
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 scripts/multi_check.py > gpurun_out/r2z2_check.log 2>&1; tail -8 gpurun_out/r2z2_check.log | cut -c1-400
timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2z2_bench_weak.json 2> gpurun_out/r2z2_bench_weak.err; head -c 330 gpurun_out/r2z2_bench_weak.json; echo; tail -2 gpurun_out/r2z2_bench_weak.err
timeout 600 $TR --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --quick > gpurun_out/r2z2_bench_strong.json 2> gpurun_out/r2z2_bench_strong.err; head -c 330 gpurun_out/r2z2_bench_strong.json; echo; tail -2 gpurun_out/r2z2_bench_strong.err
timeout 600 $TR --master-port 29614 bench.py --gpus 2 --steps 20 --warmup 5 --comm library --quick > gpurun_out/r2z2_bench_weak_libcomm.json 2> gpurun_out/r2z2_bench_weak_libcomm.err; head -c 330 gpurun_out/r2z2_bench_weak_libcomm.json; echo; tail -2 gpurun_out/r2z2_bench_weak_libcomm.err
