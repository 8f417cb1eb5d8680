#!/bin/bash
mkdir -p gpurun_out
for v in "" "--row-cost 20"; do
  tag=$(echo "rc$v" | tr -c 'a-zA-Z0-9\n' '_')
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras $v > gpurun_out/r02_n8_$tag.log 2>&1
  python - "$v" gpurun_out/r02_n8_$tag.log <<'PY'
import json, sys
ok = False
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d = json.loads(l); x = d["extra"]; ok = True; p = d["config"]["partition"]
        print(repr(sys.argv[1]), "ms_per_pass", round(d["ms_per_step"], 2), "step_alone", round(x["spmm_step_ms_alone"], 3), "xfer_alone", round(x["transfers_ms_alone"], 3),
              "e2e_ms", round(d["e2e"]["ms_per_step"], 1), "parity", d["parity"]["ok"], "rows", p["rows"], "nnz", p["nnz"], "sent", p["sent_rows"])
if not ok:
    print(repr(sys.argv[1]), "FAILED", open(sys.argv[2]).read()[-800:])
PY
done
