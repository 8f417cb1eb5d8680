#!/bin/bash
# rows per CTA of the dense-Pi build step: variants of the same library (tools/build_variants.py pprrows*)
mkdir -p gpurun_out
cat > /tmp/ppr_time.py <<'PY'
import os, sys, time, torch
sys.path.insert(0, ".")
import ppnp_b200 as P
from ppnp_b200.synth import powerlaw_adjacency
dev = torch.device("cuda:0")
n = 19717
ip, idx = powerlaw_adjacency(n, 88648, seed=0, device=dev)
ahat = P.csr_normalize(ip, idx)
for method, K in (("chebyshev", P.ppr_cheb_steps_for_tol(0.1, 1e-7)), ("power", 153)):
    P.ppr_dense(ahat, 0.1, K=K, method=method); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); Pi = P.ppr_dense(ahat, 0.1, K=K, method=method); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(os.environ.get("PPNP_B200_LIB", "default (8 rows)").split("_")[-1], method, "K", K, "ms", round(ms, 2), "ms/step", round(ms / K, 3),
          "frac of HBM peak", round(3 * n * n * 4 * K / 1e9 / (ms * 1e-3) / 6553, 3), "checksum", float(Pi.sum()), flush=True)
    del Pi
PY
python /tmp/ppr_time.py
for v in pprrows1 pprrows4 pprrows16 pprrows32; do
  PPNP_B200_LIB=$PWD/ppnp_b200/variants/libppnp_b200_$v.so python /tmp/ppr_time.py
done
timeout 600 python -m pytest tests/test_gpu_pubmed_shape.py tests/test_zz_gpu_ppr_cheb.py tests/test_gpu_parity.py -m gpu -x -q --tb=short -p no:cacheprovider -k "ppr or pi_rows or cheb or pubmed" > gpurun_out/r02_pytest_ppr.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_ppr.log; tail -4 gpurun_out/r02_pytest_ppr.log
