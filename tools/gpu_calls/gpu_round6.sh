#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -6 gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.log 2>&1; tail -c 2600 gpurun_out/bench_default.log
timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 1 --workload rmat100m > gpurun_out/bench_n1_100m.log 2>&1; tail -c 1800 gpurun_out/bench_n1_100m.log
