#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_tail.py tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_tail.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_tail.log; tail -12 gpurun_out/r02_pytest_tail.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload rmat16m --steps 3 --warmup 3 ) > gpurun_out/r02_scale_n2_16m.log 2>&1
tail -c 1800 gpurun_out/r02_scale_n2_16m.log
