#!/bin/bash
# the benched configuration (row-major stream, degree order, idx16): launch list + one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/r02_ncu_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_stream_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/r02_prof_default -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/r02_ncu_default.log 2>&1
ls -la gpurun_out/r02_prof_default.ncu-rep gpurun_out/r02_launches.csv
