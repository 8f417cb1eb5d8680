#!/bin/bash
# usage: gpu_scale2.sh N rowgroups...
mkdir -p gpurun_out
N=$1; shift
for rg in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --workload rmat100m --transport pipe --row-groups $rg > gpurun_out/bench_n${N}_100m_pipe_rg$rg.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_100m_pipe_rg$rg.log") if l.startswith("{")][-1])
    p=d["config"]["partition"]
    print("N=$N pipe rg=$rg: ms/pass %.2f value %.3e  transfers %.2f ms  spmm %.2f ms  e2e %.3e" % (d["ms_per_step"], d["value"], d["extra"]["transfers_ms_alone"], d["extra"]["spmm_step_ms_alone"], d["e2e"]["value"]))
    print("  rows", [round(x/1e6,1) for x in p["rows"]], "sent", [round(x/1e6,1) for x in p["sent_rows"]], "halo", [round(x/1e6,1) for x in p["halo_rows"]])
except Exception as e:
    print("N=$N rg=$rg: FAILED", e)
    import subprocess; print(subprocess.run("tail -5 gpurun_out/bench_n${N}_100m_pipe_rg$rg.log", shell=True, capture_output=True, text=True).stdout[-1500:])
PY
done
