#!/bin/bash
# config-5 kernel (F = 16) on the 16 M-node scale model, one GPU, partitioned code path: stream orders
mkdir -p gpurun_out
: > gpurun_out/r02_window16.jsonl
for v in "--rows-below 64" "--rows-below 64 --order window --window-key mid" "--rows-below 64 --order window --window-key first" "--rows-below 0 --order window --window-key mid" "--rows-below 0 --order carve --carve-block-cols 750000 --carve-blocks 16 --carve-min-piece 16" "--rows-below 0"; do
  timeout 400 python bench.py --gpus 1 --workload rmat16m --steps 3 --warmup 2 --no-extras --no-cpu-baseline $v > gpurun_out/r02_window16_one.log 2>&1
  grep '^{' gpurun_out/r02_window16_one.log >> gpurun_out/r02_window16.jsonl
  python - "$v" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/r02_window16_one.log") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print(repr(sys.argv[1]), "ms/pass", round(d["ms_per_step"], 2), "parity", d["parity"].get("ok"), d["config"]["partition"]["transport"], "build_s", d["extra"].get("build_s"))
else:
    print(repr(sys.argv[1]), "FAILED", open("gpurun_out/r02_window16_one.log").read()[-1200:])
PY
done
