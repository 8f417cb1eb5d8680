"""One variant of tools/bench_tiled.py, a few launches of the hub kernel and of the rest kernel alone (for ncu).
python tools/prof_tiled.py '{"slice_width": 64}' [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppnp_b200 as P  # noqa: E402
from ppnp_b200 import _lib  # noqa: E402
from ppnp_b200.synth import rmat_adjacency  # noqa: E402


def main():
    kw = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {"slice_width": 64}
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    n, raw, scale, F = 2_000_000, 26_400_000, 21, 64
    ip, idx = rmat_adjacency(n, raw, scale, seed=0, device=dev)
    ahat = P.csr_normalize(ip, idx)
    g = P.PropagationGraph(ahat, chunk_edges=256, order="degree", idx16=True, tiled=kw)
    tp, rest, W, rows = g.tiled_for(F)
    H = torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    Z = torch.empty_like(H)
    lib = _lib.load()
    for _ in range(reps):
        _lib.check(lib.ppnp_spmm_step_tiled(tp.struct(), _lib.ptr(H), _lib.ptr(H), _lib.ptr(Z), F, F, W, 0.1, _lib.EPI_Y, 0,
                                            _lib.current_stream()), "tiled")
        if rest is not None:
            _lib.check(lib.ppnp_spmm_step(rest.struct(), _lib.ptr(H), _lib.ptr(H), _lib.ptr(Z), _lib.ptr(g.rest_partial_buffer(rest, F)),
                                          F, F, 0.1, _lib.EPI_Y, 0, _lib.current_stream()), "rest")
    torch.cuda.synchronize()
    print(json.dumps(tp.stats))


if __name__ == "__main__":
    main()
