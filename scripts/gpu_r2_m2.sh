#!/bin/bash
# 2-GPU call: partitioned correctness (NCCL, graph) + weak / strong scaling bench lines
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/multi_check.py > gpurun_out/r2m2_check.log 2>&1; tail -6 gpurun_out/r2m2_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2m2_bench_weak.json 2> gpurun_out/r2m2_bench_weak.err; head -c 400 gpurun_out/r2m2_bench_weak.json; tail -2 gpurun_out/r2m2_bench_weak.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --quick > gpurun_out/r2m2_bench_strong.json 2> gpurun_out/r2m2_bench_strong.err; head -c 400 gpurun_out/r2m2_bench_strong.json; tail -2 gpurun_out/r2m2_bench_strong.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 20 --warmup 5 --no-graph --quick > gpurun_out/r2m2_bench_weak_eager.json 2> gpurun_out/r2m2_bench_weak_eager.err; head -c 300 gpurun_out/r2m2_bench_weak_eager.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg4 --quick > gpurun_out/r2m2_bench_cfg4.json 2> gpurun_out/r2m2_bench_cfg4.err; head -c 300 gpurun_out/r2m2_bench_cfg4.json
