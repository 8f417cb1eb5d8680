#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
for mode in exact appnp; do
  timeout 600 python tools/train_cora.py --mode $mode --runs 4 > gpurun_out/train_cora_$mode.jsonl 2> gpurun_out/train_cora_$mode.err; tail -1 gpurun_out/train_cora_$mode.jsonl; tail -2 gpurun_out/train_cora_$mode.err
done
timeout 600 python tools/train_cora.py --mode exact --gemm bf16 --runs 2 > gpurun_out/train_cora_bf16.jsonl 2> gpurun_out/train_cora_bf16.err; tail -1 gpurun_out/train_cora_bf16.jsonl
timeout 600 python tools/bench_exact.py > gpurun_out/bench_exact.jsonl 2> gpurun_out/bench_exact.err; grep '"N": 64' gpurun_out/bench_exact.jsonl | head -2 | cut -c1-400
timeout 600 python bench.py --gpus 1 --steps 1 --warmup 1 --workload rmat16m > gpurun_out/plain16.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 6 -c 1 -o gpurun_out/prof_spmm_f16 python bench.py --gpus 1 --steps 1 --warmup 1 --workload rmat16m > gpurun_out/ncu_f16.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
