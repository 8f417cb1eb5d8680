"""Lean check of an experimental library build (PPNP_B200_LIB=...): parity against the C oracle on a
skewed 50 k-node R-MAT graph, then ms per step on config 4 for a few plans.  Appends JSON lines to
gpurun_out/fewends.jsonl as it goes (the GPU call may be cut short)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ppnp_b200 as P  # noqa: E402
import ppnp_oracle as oracle  # noqa: E402  (checker only)
from ppnp_b200.synth import rmat_adjacency  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "fewends.jsonl")
dev = torch.device("cuda:0")


def emit(rec):
    rec["lib"] = os.path.basename(os.environ.get("PPNP_B200_LIB", "default"))
    line = json.dumps(rec)
    print(line, flush=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def main():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # ---- parity
    ip, idx = oracle.rmat_graph(50000, 1200000, 16, seed=0)
    oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
    ahat = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev), torch.from_numpy(idx).to(dev))
    plans = {"degree": dict(order="degree"), "degree+idx16": dict(order="degree", idx16=True),
             "carve128x16+idx16": dict(order="carve", idx16=True, carve=dict(block_cols=128, n_blocks=16, min_piece=3))}
    for F in (64, 16):
        Hn = np.random.RandomState(F).randn(50000, F).astype(np.float32)
        ref = oracle.c_appnp_f64(oip, oidx, oval, Hn.astype(np.float64), 10, 0.1)
        for name, kw in plans.items():
            g = P.PropagationGraph(ahat, chunk_edges=256, **kw)
            for use_vals in (False, True):
                Z = P.appnp_propagate(g, torch.from_numpy(Hn).to(dev), K=10, alpha=0.1, use_vals=use_vals).cpu().numpy()
                err = float(np.linalg.norm(Z - ref) / np.linalg.norm(ref))
                emit({"check": "parity", "plan": name, "F": F, "use_vals": use_vals, "relerr": err, "ok": bool(err < 1e-5)})
    # ---- config 4 timing
    n, raw, scale, F, K = 2_000_000, 26_400_000, 21, 64, 10
    ip, idx = rmat_adjacency(n, raw, scale, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    H = torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    Z, S = torch.empty_like(H), torch.empty_like(H)
    for name, kw in [("degree+idx16", dict(order="degree", idx16=True)),
                     ("carve512x64T4+idx16", dict(order="carve", idx16=True, carve=dict(block_cols=512, n_blocks=64, min_piece=4))),
                     ("carve512x64T8+idx16", dict(order="carve", idx16=True, carve=dict(block_cols=512, n_blocks=64, min_piece=8))),
                     ("carve512x256T3+idx16", dict(order="carve", idx16=True, carve=dict(block_cols=512, n_blocks=256, min_piece=3))),
                     ("degree", dict(order="degree"))]:
        g = P.PropagationGraph(ahat, chunk_edges=256, **kw)
        P.appnp_propagate(g, H, K, 0.1, out=Z, scratch=S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            P.appnp_propagate(g, H, K, 0.1, out=Z, scratch=S)
        e1.record()
        torch.cuda.synchronize()
        emit({"check": "time", "plan": name, "ms_per_step": e0.elapsed_time(e1) / (3 * K), "checksum": float(Z.double().sum())})
        del g


if __name__ == "__main__":
    main()
