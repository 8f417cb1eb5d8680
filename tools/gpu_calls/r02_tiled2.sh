#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_tiled.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_tiled.log; tail -5 gpurun_out/r02_pytest_tiled.log
rm -f gpurun_out/bench_tiled.jsonl
timeout 600 python tools/bench_tiled.py rowmajor+idx16 w64 w32 w64-nopace w64-slack3 w32-slack3 w64-fine128 > gpurun_out/r02_bench_tiled2.log 2>&1; cut -c1-1300 gpurun_out/r02_bench_tiled2.log
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section SchedulerStats --section WarpStateStats --section Occupancy --section LaunchStats --section InstructionStats"
timeout 300 ncu $SEC --clock-control none -k regex:spmm_tiled --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_tiled_w64_nopace -f python tools/prof_tiled.py '{"slice_width": 64, "slack": 1048576}' > gpurun_out/r02_ncu_t1.log 2>&1
timeout 300 ncu $SEC --clock-control none -k regex:spmm_tiled --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_tiled_w64 -f python tools/prof_tiled.py '{"slice_width": 64}' > gpurun_out/r02_ncu_t2.log 2>&1
timeout 300 ncu $SEC --clock-control none -k regex:spmm_stream --launch-skip 1 --launch-count 1 -o gpurun_out/r02_prof_rest_w64 -f python tools/prof_tiled.py '{"slice_width": 64}' > gpurun_out/r02_ncu_t3.log 2>&1
ls -la gpurun_out/*.ncu-rep
