#!/bin/bash
# First GPU pass: parity tests, bench variants, ncu launch list + one full capture of the SpMM kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
timeout 300 python bench.py --steps 5 --warmup 3 --order degree --no-cpu-baseline > gpurun_out/bench_degree.log 2>&1; tail -1 gpurun_out/bench_degree.log
timeout 300 python bench.py --steps 5 --warmup 3 --use-vals --no-cpu-baseline > gpurun_out/bench_vals.log 2>&1; tail -1 gpurun_out/bench_vals.log
timeout 300 python bench.py --steps 5 --warmup 3 --chunk-edges 512 --no-cpu-baseline > gpurun_out/bench_c512.log 2>&1; tail -1 gpurun_out/bench_c512.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 25 -c 2 -o gpurun_out/prof_spmm python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
