"""One process, one graph (config 4, R-MAT 2 M nodes / F = 64), several plan-selected kernel variants:
ms per propagation step (CUDA events, K = 10 forward per pass) and the difference to the default
path's result.  Appends one JSON line per variant to gpurun_out/bench_variants.jsonl as it goes."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppnp_b200 as P  # noqa: E402
from ppnp_b200.synth import rmat_adjacency  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "bench_variants.jsonl")
VARIANTS = [
    ("degree", dict(order="degree")),
    ("carve512x64-narrow", dict(order="carve", carve=dict(block_cols=512, n_blocks=64, min_piece=4, wide_cta=False))),
    ("carve512x64", dict(order="carve", carve=dict(block_cols=512, n_blocks=64, min_piece=4))),
    ("degree+idx16", dict(order="degree", idx16=True)),
    ("carve512x64+idx16", dict(order="carve", idx16=True, carve=dict(block_cols=512, n_blocks=64, min_piece=4))),
    ("carve512x256+idx16", dict(order="carve", idx16=True, carve=dict(block_cols=512, n_blocks=256, min_piece=4))),
    ("carve384x96T6+idx16", dict(order="carve", idx16=True, carve=dict(block_cols=384, n_blocks=96, min_piece=6))),
    # not measured in round 1 (GPU budget spent): carved and residual chunk units alternating
    ("carve512x64T8+idx16+interleave", dict(order="carve", idx16=True, carve=dict(block_cols=512, n_blocks=64, min_piece=8, interleave=True))),
    ("carve2level+idx16+interleave", dict(order="carve", idx16=True, carve=dict(levels=[(512, 64, 8), (125000, 16, 16)], interleave=True))),
    ("carve2level+idx16", dict(order="carve", idx16=True, carve=dict(levels=[(512, 64, 8), (125000, 16, 16)]))),
    ("carveL2only+idx16", dict(order="carve", idx16=True, carve=dict(levels=[(125000, 16, 16)], wide_cta=False))),
    ("carve512x64T4+idx16+interleave", dict(order="carve", idx16=True, carve=dict(block_cols=512, n_blocks=64, min_piece=4, interleave=True))),
]


def main():
    only = sys.argv[1:]
    dev = torch.device("cuda:0")
    n, raw, scale, F, K = 2_000_000, 26_400_000, 21, 64, 10
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    ip, idx = rmat_adjacency(n, raw, scale, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    H = torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    Z, S = torch.empty_like(H), torch.empty_like(H)
    base = None
    for name, kw in VARIANTS:
        if only and name not in only:
            continue
        rec = {"variant": name, "nnz": ahat.nnz, "F": F, "K": K}
        try:
            t0 = time.perf_counter()
            g = P.PropagationGraph(ahat, chunk_edges=256, **kw)
            torch.cuda.synchronize()
            rec["plan_s"] = round(time.perf_counter() - t0, 2)
            rec["carve"] = g.plan.carve
            rec["n_slots"], rec["n_fix"] = g.plan.n_slots, g.plan.n_fix
            for _ in range(2):
                P.appnp_propagate(g, H, K, 0.1, out=Z, scratch=S)
            torch.cuda.synchronize()
            reps = 4
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
            ev[0].record()
            for i in range(reps):
                P.appnp_propagate(g, H, K, 0.1, out=Z, scratch=S)
                ev[i + 1].record()
            torch.cuda.synchronize()
            per = [ev[i].elapsed_time(ev[i + 1]) / K for i in range(reps)]
            rec["ms_per_step"] = sum(per) / reps
            rec["ms_per_step_min"] = min(per)
            rec["edge_feature_per_s"] = ahat.nnz * F / (rec["ms_per_step"] * 1e-3)
            if base is None:
                base = Z.clone()
            else:
                rec["rel_diff_vs_default"] = float((Z - base).norm() / base.norm())
            del g
        except Exception as e:  # noqa: BLE001  (a failed variant must not hide the others' numbers)
            rec["error"] = repr(e)[:400]
        line = json.dumps(rec)
        print(line, flush=True)
        with open(OUT, "a") as f:
            f.write(line + "\n")
        if "error" in rec and "CUDA" in rec["error"]:
            break


if __name__ == "__main__":
    main()
