#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --workload rmat16m > gpurun_out/bench_n${N}_16m.log 2>&1; tail -c 2200 gpurun_out/bench_n${N}_16m.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --workload rmat100m > gpurun_out/bench_n${N}_100m.log 2>&1; tail -c 2500 gpurun_out/bench_n${N}_100m.log
