#!/bin/bash
# counters of the F = 16 kernels on the 16 M-node model, rows of a block in the generator's order and in degree order
mkdir -p gpurun_out
for v in "" "--no-degree-sort"; do
  tag=$(echo "f16$v" | tr -c 'a-zA-Z0-9\n' '_')
  timeout 400 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --clock-control none -k regex:"spmm_stream_kernel|spmm_rows_kernel" --launch-skip 12 --launch-count 2 -o gpurun_out/r02_prof_$tag -f python bench.py --gpus 1 --workload rmat16m --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras $v > gpurun_out/r02_ncu_$tag.log 2>&1
  ls -la gpurun_out/r02_prof_$tag.ncu-rep
done
