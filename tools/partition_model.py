"""DESIGN TOOL (not product code): halo volume of the 1-D row partition under three labellings of the config-4
R-MAT scale model: (a) block-cyclic stripes + nnz-balanced cuts (what ppnp_b200/dist.py does),
(b) the generator's natural ids + nnz-balanced contiguous cuts (keeps R-MAT's id-prefix locality, loses balance of the
SENT volume), (c) natural ids with the k highest-degree rows replicated on every rank (their Z rows are recomputed
locally from partial sums -- counted as k partial rows received per rank instead of halo rows).
Prints rows received / sent per rank and step and the nnz balance.
(Since the end of round 2 dist.py also weighs rows in the cut -- auto_row_cost, which moves the boundaries by a few per
cent -- and stores the rows of a block in degree order, which changes no block's membership and hence no halo volume.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ppnp_oracle as oracle  # noqa: E402
from ppnp_b200.dist import auto_stripes, balanced_row_blocks, stripe_relabel  # noqa: E402


def volumes(r, c, deg, P, n, skip_cols=None):
    bounds = np.array(balanced_row_blocks(torch.from_numpy(deg), P))
    owner = np.searchsorted(bounds, np.arange(n), side="right") - 1
    ro, co = owner[r], owner[c]
    remote = ro != co
    if skip_cols is not None:
        remote &= ~skip_cols[c]
    pairs = np.unique(ro[remote] * n + c[remote])
    recv = np.bincount(pairs // n, minlength=P)
    sent = np.bincount(owner[pairs % n], minlength=P)
    nnz = np.bincount(ro, minlength=P) + np.bincount(owner, minlength=P)
    return recv, sent, nnz


def main():
    n, raw, scale = 2_000_000, 26_400_000, 21
    ip, idx = oracle.rmat_graph(n, raw, scale)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ip))
    cols = idx.astype(np.int64)
    deg0 = np.diff(ip) + 1
    for P in (4, 8):
        st = auto_stripes(n, P)
        f = stripe_relabel(torch.arange(n), n, P, st).numpy()
        deg = np.empty(n, dtype=np.int64); deg[f] = deg0
        recv, sent, nnz = volumes(f[rows], f[cols], deg, P, n)
        print(f"P={P} (a) block-cyclic ({st} stripes): recv max {recv.max()/1e3:.0f} k mean {recv.mean()/1e3:.0f} k | sent max {sent.max()/1e3:.0f} k | "
              f"nnz max/mean {nnz.max()/nnz.mean():.3f}", flush=True)
        recv, sent, nnz = volumes(rows, cols, deg0, P, n)
        print(f"P={P} (b) natural ids:                recv max {recv.max()/1e3:.0f} k mean {recv.mean()/1e3:.0f} k | sent max {sent.max()/1e3:.0f} k | "
              f"nnz max/mean {nnz.max()/nnz.mean():.3f}", flush=True)
        order = np.argsort(-deg0, kind="stable")
        for k in (1000, 10000, 50000):
            hub = np.zeros(n, dtype=bool); hub[order[:k]] = True
            recv, sent, nnz = volumes(rows, cols, deg0, P, n, skip_cols=hub)
            print(f"P={P} (c) natural ids, top {k:6d} rows replicated: recv max {(recv.max()+k)/1e3:.0f} k (halo {recv.max()/1e3:.0f} k + {k/1e3:.0f} k partial rows) "
                  f"| sent max {(sent.max()+k)/1e3:.0f} k", flush=True)


if __name__ == "__main__":
    main()
