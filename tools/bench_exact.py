#!/usr/bin/env python
"""BASELINE.json config 2: exact PPNP on a PubMed-shape graph (n = 19 717) on one B200.
(i) Pi build: GPU power iteration vs the reference's dense fp64 inverse on the host (bounded
sample); (ii) Pi apply: fp32 SIMT and bf16 tcgen05 gather-GEMM for N in {3, 7, 16, 64}, full Pi
and the main.py row counts, as bytes of Pi streamed per second against the measured HBM peak.
One JSON object per line on stdout."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppnp_b200 as P  # noqa: E402
from ppnp_b200.synth import powerlaw_adjacency  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    n, alpha = 19717, 0.1
    ip, idx = powerlaw_adjacency(n, 88648, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    K = P.ppr_steps_for_tol(alpha, 1e-7)
    t_build = timed(lambda: P.ppr_dense(ahat, alpha, K=K), reps=3, warm=1)
    Kc = P.ppr_cheb_steps_for_tol(alpha, 1e-7)
    t_cheb = timed(lambda: P.ppr_dense(ahat, alpha, K=Kc, method="chebyshev"), reps=3, warm=1)
    Pc = P.ppr_dense(ahat, alpha, K=Kc, method="chebyshev")
    Pi = P.ppr_dense(ahat, alpha, K=K)
    out = {"what": "ppr_build", "n": n, "nnz_a": int(ip[-1]), "K": K, "ms": t_build,
           "algorithmic_GB": 3 * n * n * 4 * K / 1e9, "achieved_GBps": 3 * n * n * 4 * K / 1e9 / (t_build * 1e-3),
           "frac_of_hbm_peak": 3 * n * n * 4 * K / 1e9 / (t_build * 1e-3) / PEAK,
           "chebyshev": {"K": Kc, "ms": t_cheb, "rel_diff_vs_power": float((Pc - Pi).norm() / Pi.norm())}}
    del Pc
    # CPU reference sample: helpers.py:68-71 on the first ns nodes' induced subgraph is not the same
    # matrix; time the dense fp64 inverse itself at a bounded size and quote the n^3 scaling
    ns = int(os.environ.get("PPNP_CPU_INV_N", "5000"))
    A = np.random.RandomState(0).rand(ns, ns) * 0.01 + np.eye(ns)
    t0 = time.perf_counter(); np.linalg.inv(A); t_inv = time.perf_counter() - t0
    out["cpu_inverse_sample"] = {"n": ns, "seconds": t_inv, "extrapolated_seconds_at_n": t_inv * (n / ns) ** 3,
                                 "threads": os.cpu_count()}
    print(json.dumps(out), flush=True)

    Pb = P.to_bf16_padded(Pi)
    g = torch.Generator(device=dev).manual_seed(1)
    for N in (3, 7, 16, 64):
        H = torch.randn(n, N, device=dev, generator=g)
        for m in (None, 60, 500, 940):
            idxr = None if m is None else torch.randperm(n, device=dev, generator=g)[:m]
            rows = n if m is None else m
            t32 = timed(lambda: P.gather_gemm(Pi, H, idxr))
            t16 = timed(lambda: P.gather_gemm_bf16(Pb, H, idxr))
            ttorch = timed(lambda: (Pi if idxr is None else Pi[idxr]) @ H)     # the reference's own GPU path (model.py:63)
            ref = ((Pi if idxr is None else Pi[idxr]).double() @ H.double())
            e32 = float((P.gather_gemm(Pi, H, idxr).double() - ref).norm() / ref.norm())
            e16 = float((P.gather_gemm_bf16(Pb, H, idxr).double() - ref).norm() / ref.norm())
            print(json.dumps({"what": "ppr_apply", "N": N, "rows": rows,
                              "fp32_ms": t32, "fp32_GBps": rows * n * 4 / 1e9 / (t32 * 1e-3), "fp32_frac": rows * n * 4 / 1e9 / (t32 * 1e-3) / PEAK, "fp32_relerr": e32,
                              "bf16_ms": t16, "bf16_GBps": rows * n * 2 / 1e9 / (t16 * 1e-3), "bf16_frac": rows * n * 2 / 1e9 / (t16 * 1e-3) / PEAK, "bf16_relerr": e16,
                              "torch_gather_matmul_ms": ttorch}), flush=True)


if __name__ == "__main__":
    main()
