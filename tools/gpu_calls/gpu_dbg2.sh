#!/bin/bash
mkdir -p gpurun_out
for b in symm nccl sync; do
  PPNP_FUSED_BARRIER=$b timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "fused" > gpurun_out/pytest_fused_$b.log 2>&1; echo "barrier=$b rc=$?"; grep -E "AssertionError: |passed|failed" gpurun_out/pytest_fused_$b.log | tail -2
done
