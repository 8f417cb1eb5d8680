#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest_dist rc=$?" >> gpurun_out/pytest_dist.log; tail -12 gpurun_out/pytest_dist.log
bash tools/gpu_scale.sh 2 "two push"
