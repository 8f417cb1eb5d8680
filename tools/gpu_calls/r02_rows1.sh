#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rows.py tests/test_gpu_tiled.py tests/test_gpu_dropin.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_rows.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_rows.log; tail -15 gpurun_out/r02_pytest_rows.log
for rb in 16 32 64 128 256 100000000; do
  timeout 300 python bench.py --rows-below $rb --no-cpu-baseline --no-extras --steps 5 > gpurun_out/r02_bench_rows_$rb.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/r02_bench_rows_$rb.log"):
    if l.startswith("{"):
        d=json.loads(l); print("rows_below $rb", "ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d["parity"]["ok"], d["parity"]["fwd"]["rel_fro"])
PY
done
timeout 300 python bench.py --tiled --rows-below 1 --no-cpu-baseline --no-extras --steps 5 > gpurun_out/r02_bench_tiled_rows.log 2>&1; tail -c 600 gpurun_out/r02_bench_tiled_rows.log
