#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_partition.py -q -m gpu --timeout 300 -x 2>&1 | tail -15
