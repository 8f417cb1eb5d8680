#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log
run() { # name, env-lib, args...
  name=$1; lib=$2; shift 2
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
  unset PPNP_B200_LIB
}
run default_degree default --order degree
run default_natural default --order natural
run mb3u4_degree mb3_u4 --order degree
run mb3u8_degree mb3_u8 --order degree
run c128_degree default --order degree --chunk-edges 128
run c512_degree default --order degree --chunk-edges 512
run l2_degree default --order degree --workload rmatl2
run l2_natural default --order natural --workload rmatl2
run vals_degree default --order degree --use-vals
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 25 -c 1 -o gpurun_out/prof_spmm3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload rmatl2 > gpurun_out/plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 25 -c 1 -o gpurun_out/prof_spmm3_l2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload rmatl2 > gpurun_out/ncu_full_l2.log 2>&1
