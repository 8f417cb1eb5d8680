#!/bin/bash
# usage: gpu_scale3.sh N transport...
mkdir -p gpurun_out
N=$1; shift
for tr in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --workload rmat100m --transport $tr > gpurun_out/bench_n${N}_100m_$tr.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_100m_$tr.log") if l.startswith("{")][-1])
    p=d["config"]["partition"]
    print("N=$N $tr: ms/pass %.2f value %.3e  transfers %.2f ms  spmm %.2f ms  e2e %.3e" % (d["ms_per_step"], d["value"], d["extra"]["transfers_ms_alone"], d["extra"]["spmm_step_ms_alone"], d["e2e"]["value"]))
except Exception as e:
    print("N=$N $tr: FAILED", e)
    import subprocess; print(subprocess.run("tail -8 gpurun_out/bench_n${N}_100m_$tr.log", shell=True, capture_output=True, text=True).stdout[-2000:])
PY
done
