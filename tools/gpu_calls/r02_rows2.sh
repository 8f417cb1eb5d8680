#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rows.py tests/test_gpu_dropin.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_rows.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_rows.log; tail -8 gpurun_out/r02_pytest_rows.log
timeout 600 python tools/bench_rows.py 16000000 220000000 24 16 0 4 16 64 1073741824 > gpurun_out/r02_bench_rows_f16.log 2>&1; cat gpurun_out/r02_bench_rows_f16.log | tail -8
timeout 300 python tools/bench_rows.py 2000000 26400000 21 64 0 1073741824 > gpurun_out/r02_bench_rows_f64.log 2>&1; cat gpurun_out/r02_bench_rows_f64.log | tail -4
