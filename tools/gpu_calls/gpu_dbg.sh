#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/debug_fused.py > gpurun_out/debug_fused.log 2>&1
grep -E "rank|Error|error" gpurun_out/debug_fused.log | head -20
