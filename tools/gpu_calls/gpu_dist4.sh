#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest_dist rc=$?" >> gpurun_out/pytest_dist.log; tail -12 gpurun_out/pytest_dist.log
for cfg in "peer pull" "peer p2p" "two pull" "one p2p"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --workload rmat16m --phases $1 --transport $2 > gpurun_out/bench_n${N}_16m_$1_$2.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_16m_$1_$2.log") if l.startswith("{")][-1])
    print("$1/$2: ms/pass %.2f value %.3e  transfers %.2f ms  spmm %.2f ms" % (d["ms_per_step"], d["value"], d["extra"]["transfers_ms_alone"], d["extra"]["spmm_step_ms_alone"]))
except Exception as e:
    print("$1/$2: FAILED", e)
PY
done
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --workload rmat16m --no-cpu-baseline > gpurun_out/bench_n1_16m.log 2>&1; tail -c 300 gpurun_out/bench_n1_16m.log | head -c 300; echo
