#!/bin/bash
# usage: gpu_scale.sh N "phases transport" ...
mkdir -p gpurun_out
N=$1; shift
for cfg in "$@"; do
  set -- $cfg
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --workload rmat100m --phases $1 --transport $2 > gpurun_out/bench_n${N}_100m_$1_$2.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_100m_$1_$2.log") if l.startswith("{")][-1])
    print("N=$N $1/$2: ms/pass %.2f value %.3e  transfers %.2f ms  spmm %.2f ms  e2e %.3e" % (d["ms_per_step"], d["value"], d["extra"]["transfers_ms_alone"], d["extra"]["spmm_step_ms_alone"], d["e2e"]["value"]))
except Exception as e:
    print("N=$N $1/$2: FAILED", e)
    import subprocess; print(subprocess.run("tail -5 gpurun_out/bench_n${N}_100m_$1_$2.log", shell=True, capture_output=True, text=True).stdout[-1500:])
PY
done
