"""DESIGN TOOL (not product code): halo volume of the row-partitioned propagation with and without
computing hub rows where their columns live.

Graph: the config-4 R-MAT recipe (a scale model of config 5 -- same generator, same skew), P ranks,
block-cyclic relabelling + nnz-balanced contiguous cut exactly as ppnp_b200/dist.py does.

Today (1-D): rank p receives every distinct remote row its rows reference.
Hybrid: rows of degree >= D ("hubs") are not gathered at their owner; every rank sums the columns it owns
into one partial row per hub and ships that partial to the owner (P-1 partial rows per hub at most).  The
owner's halo then holds only what its NON-hub rows reference.
Prints rows received per rank and step (max over ranks, the figure that bounds the step).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ppnp_oracle as oracle  # noqa: E402
from ppnp_b200.dist import auto_stripes, balanced_row_blocks, stripe_relabel  # noqa: E402


def main():
    n, raw, scale = 2_000_000, 26_400_000, 21
    ip, idx = oracle.rmat_graph(n, raw, scale)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ip))
    cols = idx.astype(np.int64)
    for P in (2, 4, 8):
        st = auto_stripes(n, P)
        f = stripe_relabel(torch.arange(n), n, P, st).numpy()
        r, c = f[rows], f[cols]
        deg = np.bincount(r, minlength=n) + 1
        bounds = np.array(balanced_row_blocks(torch.from_numpy(deg), P))
        owner = np.searchsorted(bounds, np.arange(n), side="right") - 1
        ro, co = owner[r], owner[c]
        remote = ro != co
        # today: distinct (receiving rank, remote column) pairs
        key = ro[remote] * n + c[remote]
        halo = np.bincount(np.unique(key) // n, minlength=P)
        print(f"P={P} stripes={st}: 1-D halo rows per rank: max {halo.max() / 1e3:.0f} k  mean {halo.mean() / 1e3:.0f} k "
              f"({100 * halo.max() / (n / P):.0f}% of a rank's own rows)")
        for D in (64, 256, 1024, 4096):
            hub = deg >= D
            # halo of the non-hub rows only
            m = remote & ~hub[r]
            h2 = np.bincount(np.unique(ro[m] * n + c[m]) // n, minlength=P)
            # partial rows: one per (hub row, rank that owns at least one of its remote columns), received by the owner
            mh = remote & hub[r]
            pairs = np.unique(r[mh] * P + co[mh])
            part = np.bincount(owner[pairs // P], minlength=P)
            tot = h2 + part
            print(f"    hubs deg >= {D:5d} ({hub.sum():7d} rows, {100 * deg[hub].sum() / deg.sum():.0f}% of the non-zeros): "
                  f"halo {h2.max() / 1e3:.0f} k + partial rows {part.max() / 1e3:.0f} k = {tot.max() / 1e3:.0f} k per rank "
                  f"({100 * tot.max() / halo.max():.0f}% of today)", flush=True)


if __name__ == "__main__":
    main()
