#!/bin/bash
# last GPU seconds of round 1: new GPU tests first, then the kernel variants on config 4
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_standardize.py tests/test_gpu_variants.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1
echo "pytest_new rc=$?" >> gpurun_out/pytest_new.log
tail -5 gpurun_out/pytest_new.log
timeout 60 python tools/bench_variants.py > gpurun_out/bench_variants.log 2>&1
echo "bench_variants rc=$?" >> gpurun_out/bench_variants.log
tail -c 3000 gpurun_out/bench_variants.log
