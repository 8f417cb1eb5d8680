#!/bin/bash
# the very last GPU seconds of round 1: counters of the carved-stream step, then the older GPU test files
mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
timeout 28 ncu --clock-control none --metrics $M -k 'regex:spmm_stream_kernel|fixup_kernel' --launch-skip 4 --launch-count 4 --csv --log-file gpurun_out/ncu_carve.csv python tools/bench_variants.py carve512x64 > gpurun_out/ncu_carve.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_carve.log
timeout 40 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_shim.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_old.log 2>&1
echo "pytest_old rc=$?" >> gpurun_out/pytest_old.log
tail -3 gpurun_out/pytest_old.log
tail -c 1500 gpurun_out/ncu_carve.csv
