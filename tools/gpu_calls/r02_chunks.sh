#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_chunks.jsonl
for v in "--chunk-edges 256" "--chunk-edges 512" "--chunk-edges 384" "--chunk-edges 128" "--chunk-edges 1024" "--chunk-edges 256"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-parity $v > gpurun_out/r02_chunks_one.log 2>&1
  grep '^{' gpurun_out/r02_chunks_one.log >> gpurun_out/r02_chunks.jsonl || tail -5 gpurun_out/r02_chunks_one.log
  python - "$v" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/r02_chunks_one.log") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print(sys.argv[1], "ms/pass", round(d["ms_per_step"], 3), "launch_ms", round(d["roofline"]["launch_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
PY
done
