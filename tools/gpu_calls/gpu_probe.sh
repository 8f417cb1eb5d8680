#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -12 gpurun_out/topo.txt
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/p2p_probe.py > gpurun_out/p2p_probe.log 2>&1
grep -E "GB/s|us$|failed|via P2P|via SHM|NVLS|Connected|Channel 00" gpurun_out/p2p_probe.log | head -30
