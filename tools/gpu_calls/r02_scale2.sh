#!/bin/bash
# N = 2: the bench line of the partitioned path with the parity block, the pipelined e2e and the live T_1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rows.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_rows.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_rows.log; tail -8 gpurun_out/r02_pytest_rows.log
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dropin.log; tail -8 gpurun_out/r02_pytest_dropin.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02_scale_n2.log 2>&1
tail -c 4000 gpurun_out/r02_scale_n2.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r02_scale_n2_ref.log 2>&1
tail -c 1500 gpurun_out/r02_scale_n2_ref.log
