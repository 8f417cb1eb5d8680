#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_dist.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest_dist rc=$?" >> gpurun_out/pytest_dist.log
tail -15 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --workload rmat16m > gpurun_out/bench_n2_16m.log 2>&1; tail -c 2500 gpurun_out/bench_n2_16m.log
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --workload rmat16m --no-cpu-baseline > gpurun_out/bench_n1_16m.log 2>&1; tail -c 1500 gpurun_out/bench_n1_16m.log
