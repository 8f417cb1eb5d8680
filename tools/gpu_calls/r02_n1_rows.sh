#!/bin/bash
mkdir -p gpurun_out
for v in "--rows-below 0" "--rows-below 64"; do
  timeout 400 python bench.py --gpus 1 --workload rmat16m --steps 3 --warmup 2 --no-extras --no-cpu-baseline $v > gpurun_out/r02_n1_16m.log 2>&1
  python - "$v" <<'PY'
import json, sys
ok=False
for l in open("gpurun_out/r02_n1_16m.log"):
    if l.startswith("{"):
        d = json.loads(l); ok=True
        print(repr(sys.argv[1]), "ms_per_pass", round(d["ms_per_step"], 2), "parity", d["parity"], d["config"]["partition"]["transport"])
if not ok: print(repr(sys.argv[1]), "FAILED", open("gpurun_out/r02_n1_16m.log").read()[-1500:])
PY
done
timeout 400 python bench.py --gpus 1 --workload rmat100m --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity > gpurun_out/r02_n1_100m.log 2>&1; tail -c 700 gpurun_out/r02_n1_100m.log
bash tools/gpu_calls/r02_ncu_default.sh
