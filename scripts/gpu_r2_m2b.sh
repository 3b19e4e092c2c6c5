#!/bin/bash
# 2-GPU call (short, strict timeouts): does the graph-captured partitioned bench exit cleanly now?
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
date +%s > gpurun_out/r2m2b_t0
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2m2b_bench_weak.json 2> gpurun_out/r2m2b_bench_weak.err; echo "weak rc=$? at $(( $(date +%s) - $(cat gpurun_out/r2m2b_t0) ))s"; head -c 300 gpurun_out/r2m2b_bench_weak.json; echo
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg4 --quick > gpurun_out/r2m2b_bench_cfg4.json 2> gpurun_out/r2m2b_bench_cfg4.err; echo "cfg4 rc=$? at $(( $(date +%s) - $(cat gpurun_out/r2m2b_t0) ))s"; head -c 300 gpurun_out/r2m2b_bench_cfg4.json; echo
timeout 200 python -m pytest tests/test_gpu_partition.py -q -m gpu -k two_gpu --timeout 190 2>&1 | tail -3; echo "test at $(( $(date +%s) - $(cat gpurun_out/r2m2b_t0) ))s"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > gpurun_out/r2m2b_ref.json 2> gpurun_out/r2m2b_ref.err; echo "ref rc=$? at $(( $(date +%s) - $(cat gpurun_out/r2m2b_t0) ))s"; head -c 200 gpurun_out/r2m2b_ref.json
