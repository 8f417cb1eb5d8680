#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "fused or pipe" > gpurun_out/pytest_dist.log 2>&1; echo "pytest_dist rc=$?" >> gpurun_out/pytest_dist.log; tail -12 gpurun_out/pytest_dist.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --workload rmat100m --transport fused > gpurun_out/bench_n${N}_100m_fused.log 2>&1
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n2_100m_fused.log") if l.startswith("{")][-1])
    print("N=2 fused: ms/pass %.2f value %.3e  transfers %.2f ms  spmm %.2f ms  e2e %.3e" % (d["ms_per_step"], d["value"], d["extra"]["transfers_ms_alone"], d["extra"]["spmm_step_ms_alone"], d["e2e"]["value"]))
except Exception as e:
    print("FAILED", e); import subprocess; print(subprocess.run("tail -8 gpurun_out/bench_n2_100m_fused.log", shell=True, capture_output=True, text=True).stdout[-2000:])
PY
