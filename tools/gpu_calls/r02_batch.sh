#!/bin/bash
# config 3: the one-launch support + column map (ppnp_batch_support_colmap)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_shim.py tests/test_gpu_dropin.py -m gpu -x -q --tb=short -p no:cacheprovider -k "batch or shim or dropin" > gpurun_out/r02_pytest_batch.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_batch.log; tail -5 gpurun_out/r02_pytest_batch.log
timeout 600 python tools/bench_batch.py > gpurun_out/r02_bench_batch.jsonl 2> gpurun_out/r02_bench_batch.err; tail -3 gpurun_out/r02_bench_batch.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_batch.jsonl"):
    d = json.loads(l)
    if d["what"] == "batch":
        print(d["k"], d["B"], "ours", round(d["ours_ms"], 4), "torch dense gpu", round(d["dense_torch_gpu_ms"], 4), "cpu", round(d["dense_torch_cpu_ms"], 2), "err", d["relerr_vs_dense"])
PY
timeout 300 python bench.py --workload pubmed_batch --batch-size 32 --steps 5 --warmup 3 > gpurun_out/r02_pubmed_batch32.log 2>&1; tail -c 1500 gpurun_out/r02_pubmed_batch32.log
