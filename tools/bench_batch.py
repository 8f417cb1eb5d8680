#!/usr/bin/env python
"""BASELINE.json config 3: batch-main.py's mini-batch PPR propagation on a PubMed-shape graph, one
B200, batch-size sweep.  Per batch: support union + compaction + propagate on the compact top-k
Pi (ppnp_b200) against the literal dense lines batch-main.py:140-146 run with torch on the same GPU
and on the host.  One JSON object per line."""
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


def timed(fn, reps=30, warm=5):
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
    n, alpha, C = 19717, 0.1, 3
    ip, idx = powerlaw_adjacency(n, 88648, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    Pi0 = P.ppr_dense(ahat, alpha, tol=1e-7)
    g = torch.Generator(device=dev).manual_seed(3)
    H = torch.randn(n, C, device=dev, generator=g)
    for k in (128, 256):
        Pi = Pi0.clone()
        t_topk = timed(lambda: P.topk_thresh(Pi0, k), reps=3, warm=1)
        t_torch_topk = timed(lambda: Pi0.topk(k, -1), reps=3, warm=1)
        th = P.topk_sparsify_(Pi, k)
        assert torch.equal(th, Pi0.topk(k, -1).values[:, -1])
        spp = P.dense_to_sparse_ppr(Pi)
        print(json.dumps({"what": "topk", "k": k, "thresh_ms": t_topk, "torch_topk_ms": t_torch_topk,
                          "kept": int(spp.indices.numel()), "density": spp.indices.numel() / n / n}), flush=True)
        Pi_cpu = Pi.cpu()
        H_cpu = H.cpu()
        for B in (32, 64, 128, 256, 512, 1024):
            idx_b = torch.randperm(n, device=dev, generator=g)[:B].sort().values

            def ours():
                sel = P.batch_support(spp, idx_b)
                return P.batch_propagate(spp, idx_b, sel, H[sel])

            def dense_gpu():
                sub = Pi[idx_b]; sel = (sub > 0).any(dim=0); sub = sub[:, sel]
                return sub @ H[sel]

            t_ours, t_dense = timed(ours), timed(dense_gpu)
            ib = idx_b.cpu()
            t0 = time.perf_counter()
            for _ in range(3):
                sub = Pi_cpu[ib]; sel = (sub > 0).any(dim=0); sub = sub[:, sel]; _ = sub @ H_cpu[sel]
            t_cpu = (time.perf_counter() - t0) / 3 * 1e3
            err = float((ours().double() - dense_gpu().double()).norm() / dense_gpu().double().norm())
            nnz_rows = int((spp.indptr[idx_b + 1] - spp.indptr[idx_b]).sum())
            print(json.dumps({"what": "batch", "k": k, "B": B, "ours_ms": t_ours, "batches_per_s": 1e3 / t_ours,
                              "dense_torch_gpu_ms": t_dense, "dense_torch_cpu_ms": t_cpu, "relerr_vs_dense": err,
                              "kept_entries_in_batch": nnz_rows, "compact_bytes": nnz_rows * 8,
                              "dense_reference_bytes": 3 * B * n * 4}), flush=True)


if __name__ == "__main__":
    main()
