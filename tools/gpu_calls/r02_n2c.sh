#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short -p no:cacheprovider -k "fusedrows or x/fused" > gpurun_out/r02_pytest_dist_rows.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_dist_rows.log; tail -6 gpurun_out/r02_pytest_dist_rows.log
for v in "" "--rows-below 64" "--rows-below 64 --rows-order degree" "--rows-below 16"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 2 --no-extras --no-parity $v > gpurun_out/r02_n2_rows.log 2>&1
  python - "$v" <<'PY'
import json, sys
for l in open("gpurun_out/r02_n2_rows.log"):
    if l.startswith("{"):
        d = json.loads(l); x = d["extra"]
        print(repr(sys.argv[1]), "ms_per_pass", round(d["ms_per_step"], 2), "step_alone", round(x["spmm_step_ms_alone"], 3), "xfer_alone", round(x["transfers_ms_alone"], 3), d["config"]["partition"]["transport"])
PY
  tail -3 gpurun_out/r02_n2_rows.log | grep -i "error\|Traceback" | head -3
done
