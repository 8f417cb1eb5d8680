#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_ll.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
ls -la gpurun_out/launches_final.csv
