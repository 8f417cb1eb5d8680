#!/bin/bash
# in-kernel fix-up: parity, determinism, A/B against the two-launch form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_fixup.py tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_rows.py tests/test_gpu_tiled.py tests/test_gpu_plan_build.py tests/test_zz_sparse_input.py tests/test_gpu_fused_tail.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_fusedfix.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_fusedfix.log; tail -12 gpurun_out/r02_pytest_fusedfix.log
: > gpurun_out/r02_fusedfix.jsonl
for e in 1 0 1 0; do
  PPNP_FUSED_FIXUP=$e timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_fusedfix_one.log 2>&1
  grep '^{' gpurun_out/r02_fusedfix_one.log >> gpurun_out/r02_fusedfix.jsonl || tail -5 gpurun_out/r02_fusedfix_one.log
  python - "$e" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/r02_fusedfix_one.log") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print("PPNP_FUSED_FIXUP=" + sys.argv[1], "ms/pass", round(d["ms_per_step"], 3), "launch_ms", round(d["roofline"]["launch_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"], "parity", d["parity"]["ok"], d["parity"]["fwd"]["rel_fro"])
PY
done
for e in 1 0; do
  PPNP_FUSED_FIXUP=$e timeout 400 python bench.py --gpus 1 --workload rmat16m --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r02_fusedfix_16m.log 2>&1
  python - "$e" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/r02_fusedfix_16m.log") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print("rmat16m PPNP_FUSED_FIXUP=" + sys.argv[1], "ms/pass", round(d["ms_per_step"], 2), "parity", d["parity"]["ok"], "launches", d["gpu_launches"])
else:
    print(open("gpurun_out/r02_fusedfix_16m.log").read()[-1500:])
PY
done
