#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
run() { name=$1; lib=$2; shift 2
  if [ "$lib" = default ]; then unset PPNP_B200_LIB; else export PPNP_B200_LIB=$PWD/ppnp_b200/variants/libppnp_b200_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1])
    print("$name: ms/pass %.2f  frac %.3f  value %.3e e2e %.3e" % (d["ms_per_step"], d["roofline"]["frac"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$name: FAILED", e)
PY
  unset PPNP_B200_LIB; }
run stage_degree default --order degree
run stage_natural default --order natural
run stage_mb3 mb3_u4 --order degree
run stage_vals default --order degree --use-vals
run stage_l2 default --order degree --workload rmatl2
run stage_16m default --order degree --workload rmat16m
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 25 -c 1 -o gpurun_out/prof_spmm5 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 python tools/bench_exact.py > gpurun_out/bench_exact.jsonl 2> gpurun_out/bench_exact.err; tail -3 gpurun_out/bench_exact.err; head -c 1500 gpurun_out/bench_exact.jsonl
timeout 600 python tools/bench_batch.py > gpurun_out/bench_batch.jsonl 2> gpurun_out/bench_batch.err; tail -3 gpurun_out/bench_batch.err; head -c 800 gpurun_out/bench_batch.jsonl
