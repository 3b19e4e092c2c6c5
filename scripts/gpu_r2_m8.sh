#!/bin/bash
# 8-GPU call (short, strict timeouts): weak and strong scaling lines of the graph-captured partitioned path, cfg4 over 8 GPUs
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
date +%s > gpurun_out/r2m8_t0
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 5 --quick > gpurun_out/r2m8_bench_weak.json 2> gpurun_out/r2m8_bench_weak.err; echo "weak rc=$? at $(( $(date +%s) - $(cat gpurun_out/r2m8_t0) ))s"; head -c 300 gpurun_out/r2m8_bench_weak.json; echo
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 8 --steps 20 --warmup 5 --scaling strong --quick > gpurun_out/r2m8_bench_strong.json 2> gpurun_out/r2m8_bench_strong.err; echo "strong rc=$? at $(( $(date +%s) - $(cat gpurun_out/r2m8_t0) ))s"; head -c 300 gpurun_out/r2m8_bench_strong.json; echo
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 8 --steps 20 --warmup 5 --workload cfg4 --quick > gpurun_out/r2m8_bench_cfg4.json 2> gpurun_out/r2m8_bench_cfg4.err; echo "cfg4 rc=$? at $(( $(date +%s) - $(cat gpurun_out/r2m8_t0) ))s"; head -c 300 gpurun_out/r2m8_bench_cfg4.json; echo
