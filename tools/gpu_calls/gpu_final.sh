#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log; tail -4 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; tail -c 2700 gpurun_out/bench_final.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_final.log 2>&1; tail -c 900 gpurun_out/bench_ref_final.log
