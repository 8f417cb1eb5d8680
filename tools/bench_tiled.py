"""Config 4 (R-MAT 2 M nodes, F = 64): the hub rows through the shared-memory-resident kernel (csrc/appnp_tiled.cu),
the rest through the row-major stream, for several plan parameters.  ms per propagation step (CUDA events over
K = 10 forward passes) next to the row-major default, difference of the results, plan statistics.
One JSON line per variant -> gpurun_out/bench_tiled.jsonl.   python tools/bench_tiled.py [variant ...]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppnp_b200 as P  # noqa: E402
from ppnp_b200.synth import rmat_adjacency  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "bench_tiled.jsonl")
VARIANTS = [
    ("rowmajor+idx16", None),
    ("w64", dict(slice_width=64)),
    ("w32", dict(slice_width=32)),
    ("w16", dict(slice_width=16)),
    ("w64-fine128", dict(slice_width=64, fine_cols=128)),
    ("w64-slack3", dict(slice_width=64, slack=3)),
    ("w64-nopace", dict(slice_width=64, slack=1 << 20)),
    ("w32-fine512", dict(slice_width=32, fine_cols=512)),
    ("w32-slack3", dict(slice_width=32, slack=3)),
    ("w32-w8", dict(slice_width=32, warps_per_cta=8)),
    ("w64-w8", dict(slice_width=64, warps_per_cta=8)),
    ("w64-deg128", dict(slice_width=64, min_hub_degree=128)),
    ("w32-reuse1", dict(slice_width=32, fine_min_reuse=1.0)),
    ("w32-reuse3", dict(slice_width=32, fine_min_reuse=3.0)),
]


def timed(fn, reps=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]


def main():
    only = sys.argv[1:]
    dev = torch.device("cuda:0")
    n, raw, scale, F, K = 2_000_000, 26_400_000, 21, 64, 10
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    ip, idx = rmat_adjacency(n, raw, scale, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    H = torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    Z, S = torch.empty_like(H), torch.empty_like(H)
    g0 = P.PropagationGraph(ahat, chunk_edges=256, order="degree", idx16=True)
    base = P.appnp_propagate(g0, H, K, 0.1).clone()
    for name, kw in VARIANTS:
        if only and name not in only:
            continue
        rec = {"variant": name, "nnz": ahat.nnz, "F": F, "K": K}
        try:
            t0 = time.perf_counter()
            g = g0 if kw is None else P.PropagationGraph(ahat, chunk_edges=256, order="degree", idx16=True, tiled=kw)
            if kw is not None:
                tp, rest, W, rows = g.tiled_for(F)
                rec["stats"] = tp.stats
                rec["rest_edges"] = None if rest is None else rest.nnz
            torch.cuda.synchronize()
            rec["plan_s"] = round(time.perf_counter() - t0, 2)
            ms = timed(lambda: P.appnp_propagate(g, H, K, 0.1, out=Z, scratch=S))
            rec["ms_per_step"] = sum(ms) / len(ms) / K
            rec["ms_per_step_min"] = min(ms) / K
            rec["rel_diff_vs_rowmajor"] = float((Z - base).norm() / base.norm())
            if kw is not None:
                # the two kernels of a step, each alone (value-free middle step)
                from ppnp_b200 import _lib
                lib = _lib.load()
                T = H
                def hub():
                    _lib.check(lib.ppnp_spmm_step_tiled(tp.struct(), _lib.ptr(H), _lib.ptr(T), _lib.ptr(Z), F, F, W, 0.1, _lib.EPI_Y, 0,
                                                        _lib.current_stream()), "tiled")
                def rst():
                    _lib.check(lib.ppnp_spmm_step(rest.struct(), _lib.ptr(H), _lib.ptr(T), _lib.ptr(Z), _lib.ptr(g.rest_partial_buffer(rest, F)),
                                                  F, F, 0.1, _lib.EPI_Y, 0, _lib.current_stream()), "rest")
                rec["hub_kernel_ms"] = min(timed(hub, 6))
                if rest is not None:
                    rec["rest_kernels_ms"] = min(timed(rst, 6))
            del g
        except Exception as e:  # keep going: one bad variant must not lose the others
            rec["error"] = repr(e)[:400]
        print(json.dumps(rec), flush=True)
        with open(OUT, "a") as f:
            f.write(json.dumps(rec) + "\n")
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
