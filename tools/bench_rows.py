"""Row-major edge stream against the one-lane-group-per-row kernel (csrc/appnp_rows.cu) for the rows below a degree
threshold, on an R-MAT graph of the bench recipe.  ms per propagation step (CUDA events, K = 10).
python tools/bench_rows.py n raw_draws scale F [rows_below ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppnp_b200 as P  # noqa: E402
from ppnp_b200.synth import rmat_adjacency  # noqa: E402


def main():
    n, raw, scale, F = (int(x) for x in sys.argv[1:5])
    cuts = [int(x) for x in sys.argv[5:]] or [0, 8, 32, 128, 1 << 30]
    dev = torch.device("cuda:0")
    ip, idx = rmat_adjacency(n, raw, scale, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    del ip, idx
    H = torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    Z, S = torch.empty_like(H), torch.empty_like(H)
    base = None
    for rb in cuts:
        g = P.PropagationGraph(ahat, chunk_edges=256, order="degree", idx16=True, rows_below=(rb or None))
        for _ in range(2):
            P.appnp_propagate(g, H, 10, 0.1, out=Z, scratch=S)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            P.appnp_propagate(g, H, 10, 0.1, out=Z, scratch=S)
        b.record()
        torch.cuda.synchronize()
        if base is None:
            base = Z.clone()
        print(json.dumps({"n": n, "nnz": ahat.nnz, "F": F, "rows_below": rb, "ms_per_step": a.elapsed_time(b) / 30,
                          "rows_part": 0 if g.rows_part is None else int(g.rows_part[0].numel()),
                          "rel_diff": float((Z - base).norm() / base.norm())}), flush=True)
        del g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
