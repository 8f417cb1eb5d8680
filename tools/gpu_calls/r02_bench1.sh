#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dropin.log; tail -15 gpurun_out/r02_pytest_dropin.log
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default.log 2>&1; tail -c 3500 gpurun_out/r02_bench_default.log
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02_bench_ref.log 2>&1; tail -c 1500 gpurun_out/r02_bench_ref.log
( time timeout 600 python bench.py --workload pubmed_exact ) > gpurun_out/r02_bench_exact.log 2>&1; tail -c 2500 gpurun_out/r02_bench_exact.log
( time timeout 600 python bench.py --workload pubmed_batch ) > gpurun_out/r02_bench_batch.log 2>&1; tail -c 2500 gpurun_out/r02_bench_batch.log
( time timeout 600 python bench.py --impl reference --workload pubmed_exact ) > gpurun_out/r02_bench_ref_exact.log 2>&1; tail -c 800 gpurun_out/r02_bench_ref_exact.log
