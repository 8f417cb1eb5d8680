#!/bin/bash
# partitioned path: rows of a block stored in degree order (dist.degree_sort_relabel)
mkdir -p gpurun_out
: > gpurun_out/r02_degsort.jsonl
for v in "--workload rmat16m" "--workload rmat16m --no-degree-sort" "--workload rmat100m --no-parity" "--workload rmat100m --no-parity --no-degree-sort"; do
  timeout 500 python bench.py --gpus 1 $v --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r02_degsort_one.log 2>&1
  grep '^{' gpurun_out/r02_degsort_one.log >> gpurun_out/r02_degsort.jsonl
  python - "$v" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/r02_degsort_one.log") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print(repr(sys.argv[1]), "ms/pass", round(d["ms_per_step"], 2), "parity", (d["parity"] or {}).get("ok"), d["config"]["partition"]["rule"][-40:], "build_s", d["extra"].get("graph_build_s"))
else:
    print(repr(sys.argv[1]), "FAILED", open("gpurun_out/r02_degsort_one.log").read()[-1500:])
PY
done
