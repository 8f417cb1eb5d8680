#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "fused or 4/pipe" > gpurun_out/pytest_dist.log 2>&1; echo "pytest_dist rc=$?" >> gpurun_out/pytest_dist.log; tail -4 gpurun_out/pytest_dist.log
