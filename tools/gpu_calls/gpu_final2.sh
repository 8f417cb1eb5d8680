#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "fused" > gpurun_out/pytest_dist.log 2>&1; echo "pytest_dist rc=$?" >> gpurun_out/pytest_dist.log; tail -3 gpurun_out/pytest_dist.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_dist.py > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log; tail -3 gpurun_out/pytest_final.log
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; tail -c 2700 gpurun_out/bench_final.log | head -c 700
