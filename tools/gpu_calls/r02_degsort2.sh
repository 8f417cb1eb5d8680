#!/bin/bash
mkdir -p gpurun_out
for v in "--order window --window-key mid" "--order window --window-key first" "--order carve --carve-block-cols 750000 --carve-blocks 16 --carve-min-piece 16 --rows-below 0" "--rows-below 32" "--rows-below 128" "--rows-below 0"; do
  timeout 500 python bench.py --gpus 1 --workload rmat16m $v --steps 3 --warmup 2 --no-extras --no-cpu-baseline --no-parity > gpurun_out/r02_degsort2_one.log 2>&1
  python - "$v" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/r02_degsort2_one.log") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print(repr(sys.argv[1]), "ms/pass", round(d["ms_per_step"], 2), d["config"]["partition"]["transport"])
else:
    print(repr(sys.argv[1]), "FAILED", open("gpurun_out/r02_degsort2_one.log").read()[-800:])
PY
done
