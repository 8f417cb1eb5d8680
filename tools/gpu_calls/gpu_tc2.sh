#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest_tc rc=$?" >> gpurun_out/pytest_tc.log; tail -3 gpurun_out/pytest_tc.log
timeout 300 python tools/bench_exact.py > gpurun_out/bench_exact.jsonl 2> gpurun_out/bench_exact.err; python - <<'PY'
import json
for l in open("gpurun_out/bench_exact.jsonl"):
    d=json.loads(l)
    if d["what"]=="ppr_apply" and d["rows"] in (19717, 940): print(d["N"], d["rows"], "bf16 %.3f ms %.0f GB/s frac %.2f | fp32 %.3f ms | torch %.3f ms" % (d["bf16_ms"], d["bf16_GBps"], d["bf16_frac"], d["fp32_ms"], d["torch_gather_matmul_ms"]))
PY
