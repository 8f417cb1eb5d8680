#!/bin/bash
# Prepared at the end of round 1 (GPU budget spent): the first call of the next round.
#   gpurun --timeout 900 -- 'bash tools/gpu_calls/r02_first_call.sh'
# 1. everything that has not run on a GPU yet (Chebyshev Pi build, file pipeline) with the whole suite
# 2. the bench line, the exact-PPNP bench (now with the Chebyshev build), the plan variants incl. interleaved /
#    two-level carves
# 3. one full ncu capture each of the carved step and its interleaved twin (source view: where the time goes)
mkdir -p gpurun_out
PPNP_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest.log; tail -4 gpurun_out/r02_pytest.log
timeout 200 python bench.py > gpurun_out/r02_bench.log 2>&1; tail -c 1500 gpurun_out/r02_bench.log
timeout 200 python bench.py --order auto --no-cpu-baseline > gpurun_out/r02_bench_auto.log 2>&1; tail -c 1500 gpurun_out/r02_bench_auto.log
timeout 120 python tools/bench_standardize.py > gpurun_out/r02_bench_standardize.json 2> gpurun_out/r02_bench_standardize.err; cat gpurun_out/r02_bench_standardize.json
timeout 120 python tools/bench_exact.py > gpurun_out/r02_bench_exact.jsonl 2> gpurun_out/r02_bench_exact.err; head -c 900 gpurun_out/r02_bench_exact.jsonl
rm -f gpurun_out/bench_variants.jsonl
timeout 120 python tools/bench_variants.py > gpurun_out/r02_bench_variants.log 2>&1; cut -c1-220 gpurun_out/r02_bench_variants.log
for v in carve512x64+idx16 carve512x64T8+idx16+interleave; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:spmm_stream_kernel --launch-skip 3 --launch-count 1 \
      -o gpurun_out/r02_prof_${v//+/_} -f python tools/bench_variants.py $v > gpurun_out/r02_ncu_${v//+/_}.log 2>&1
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmm_stream_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02_prof_default -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_default.log 2>&1
ls -la gpurun_out | tail -8
# multi-GPU (run with gpurun --gpus 2 / 8): the hybrid exchange and the L2-carved shard streams against the fused default
#   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus N --steps 3 --warmup 2 [--dist-idx16 | --transport hybrid --hub-degree 64 | --order carve --carve-block-cols 6000000 --carve-blocks 16 --carve-min-piece 16]
#   PPNP_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q   (2 GPUs)
