#!/bin/bash
# first run of the shared-memory-resident hub-row kernel: parity, then the variants on config 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_tiled.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_tiled.log; tail -30 gpurun_out/r02_pytest_tiled.log
rm -f gpurun_out/bench_tiled.jsonl
timeout 900 python tools/bench_tiled.py > gpurun_out/r02_bench_tiled.log 2>&1; cut -c1-1200 gpurun_out/r02_bench_tiled.log
