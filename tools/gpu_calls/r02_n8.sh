#!/bin/bash
# N = 8: the fused transport and its round-2 variants on config 5 (T_1 = 776 ms measured in the N = 1 run of this round)
mkdir -p gpurun_out
: > gpurun_out/r02_n8_summary.txt
for v in "" "--rows-below 64" "--rows-below 64 --rows-order degree" "--transport hybrid --hub-degree 64" "--dist-idx16"; do
  tag=$(echo "default $v" | tr -c 'a-zA-Z0-9\n' '_')
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras $v > gpurun_out/r02_n8_$tag.log 2>&1
  python - "$v" gpurun_out/r02_n8_$tag.log <<'PY' | tee -a gpurun_out/r02_n8_summary.txt
import json, sys
ok = False
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d = json.loads(l); x = d["extra"]; ok = True
        print(repr(sys.argv[1]), "ms_per_pass", round(d["ms_per_step"], 2), "eff_vs_776", round(776.0 / 8 / d["ms_per_step"], 4), "step_alone", round(x["spmm_step_ms_alone"], 3),
              "xfer_alone", round(x["transfers_ms_alone"], 3), "e2e_ms", round(d["e2e"]["ms_per_step"], 1), "parity", d["parity"]["ok"], d["parity"].get("adjointness_rel_full_size"),
              d["config"]["partition"]["transport"])
if not ok:
    print(repr(sys.argv[1]), "FAILED", open(sys.argv[2]).read()[-600:])
PY
done
