#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log; tail -4 gpurun_out/pytest_final.log
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; tail -c 2800 gpurun_out/bench_final.log
