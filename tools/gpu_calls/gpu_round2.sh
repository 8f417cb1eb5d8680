#!/bin/bash
# Round-1 GPU pass 2: parity of the rewritten SpMM kernel, tuning variants, tcgen05 tests (last, own timeout).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for v in default mb3_u4 mb3_u8 mb2_u8 mb4_u2; do
  if [ "$v" = default ]; then unset PPNP_B200_LIB; else export PPNP_B200_LIB=$PWD/ppnp_b200/variants/libppnp_b200_$v.so; fi
  for order in natural degree; do
    timeout 300 python bench.py --steps 5 --warmup 3 --order $order --no-cpu-baseline > gpurun_out/bench_${v}_${order}.log 2>&1
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${v}_${order}.log").read().strip().splitlines()[-1])
    print("${v} ${order}: ms/pass %.2f  frac %.3f  value %.3e" % (d["ms_per_step"], d["roofline"]["frac"], d["value"]))
except Exception as e:
    print("${v} ${order}: FAILED", e)
PY
  done
done
unset PPNP_B200_LIB
timeout 300 python bench.py --steps 5 --warmup 3 --use-vals --no-cpu-baseline > gpurun_out/bench_vals.log 2>&1; tail -c 600 gpurun_out/bench_vals.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 25 -c 2 -o gpurun_out/prof_spmm2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest_tc rc=$?" >> gpurun_out/pytest_tc.log
tail -30 gpurun_out/pytest_tc.log
