#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plan_build.py -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_planbuild.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_planbuild.log; tail -15 gpurun_out/r02_pytest_planbuild.log
python - <<'PY'
import time, torch, sys
sys.path.insert(0, ".")
import ppnp_b200 as P
from ppnp_b200.synth import rmat_adjacency
from ppnp_b200.plan import build_stream_plan_cuda, build_stream_plan_torch, degree_order, lane_transpose
dev = torch.device("cuda:0")
ip, idx = rmat_adjacency(2_000_000, 26_400_000, 21, seed=0, device=dev)
ahat = P.csr_normalize(ip, idx)
o = degree_order(ahat.indptr)
for name, fn in (("cuda", lambda: build_stream_plan_cuda(ahat.indptr, ahat.indices, ahat.val32, 256, o)),
                 ("cuda G=16", lambda: build_stream_plan_cuda(ahat.indptr, ahat.indices, ahat.val32, 256, o, lane_group=16)),
                 ("torch", lambda: build_stream_plan_torch(ahat.indptr, ahat.indices, ahat.val32, 256, o)),
                 ("torch + lane_transpose", lambda: lane_transpose(build_stream_plan_torch(ahat.indptr, ahat.indices, ahat.val32, 256, o), 16))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): p = fn()
    torch.cuda.synchronize()
    print(f"config-4 plan, {name}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms", flush=True)
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_planbuild_bench.log 2>&1; python -c "
import json
for l in open('gpurun_out/r02_planbuild_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('bench ms/pass', d['ms_per_step'], 'build_s', d['config']['graph_build_s'], 'parity', d['parity']['ok'])
"
timeout 400 python bench.py --gpus 1 --workload rmat16m --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r02_planbuild_16m.log 2>&1; python -c "
import json
for l in open('gpurun_out/r02_planbuild_16m.log'):
    if l.startswith('{'):
        d=json.loads(l); print('rmat16m ms/pass', d['ms_per_step'], 'parity', d['parity']['ok'], {k:v for k,v in d['extra'].items() if 'build' in k})
" || tail -20 gpurun_out/r02_planbuild_16m.log
