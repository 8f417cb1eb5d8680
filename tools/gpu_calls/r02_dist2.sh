#!/bin/bash
# 2 GPUs: every multi-GPU transport against the C oracle, including the cases written at the end of round 1
mkdir -p gpurun_out
PPNP_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_dist2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dist2.log; tail -15 gpurun_out/r02_pytest_dist2.log
