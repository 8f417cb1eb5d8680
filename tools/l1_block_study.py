"""DESIGN TOOL (not product code): how much of the L2->SM gather traffic of one APPNP step could
an L1-resident column block remove on the config-4 R-MAT graph?

Model.  Columns are ranked by degree; block b = ranks [b*BS, (b+1)*BS).  A (row, block) *piece* is
carved out of its row when it has >= T edges; carved pieces are streamed block-major so that one
SM walks a run of edges whose columns all lie in one L1-sized block (BS rows of F*4 bytes).  A
carved piece costs one partial-sum write + one read (2 row transfers) and its edges hit L1 once
the block is resident; every other edge still pulls its row from L2.

Prints, per (BS, T, NB): carved edges, pieces, and rows moved L2->SM relative to today's nnz.
"""
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import ppnp_oracle as oracle  # noqa: E402


def main():
    n, raw, scale = 2_000_000, 26_400_000, 21
    if len(sys.argv) > 1:
        n, raw, scale = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    t = time.time()
    indptr, indices = oracle.rmat_graph(n, raw, scale)
    deg = np.diff(indptr) + 1                      # A_hat rows hold the self loop
    nnz = int(indices.size) + n
    print(f"graph {time.time() - t:.1f}s n={n} nnz(A_hat)={nnz}")
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    rows = np.concatenate([rows, np.arange(n, dtype=np.int64)])
    cols = np.concatenate([indices.astype(np.int64), np.arange(n, dtype=np.int64)])
    order = np.argsort(-deg, kind="stable")
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    crank = rank[cols]
    sd = np.sort(deg)[::-1]
    cum = np.cumsum(sd) / nnz
    for k in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072):
        print(f"  top {k:7d} columns receive {100 * cum[k - 1]:.1f}% of the gathers")
    for BS in (512, 1024, 2048):
        for NB in (16, 64, 256):
            lim = BS * NB
            m = crank < lim
            key = rows[m] * NB + crank[m] // BS
            uk, cnt = np.unique(key, return_counts=True)
            for T in (3, 4, 6, 8):
                sel = cnt >= T
                carved = int(cnt[sel].sum())
                pieces = int(sel.sum())
                # rows that would be split into >1 piece need slots; a row entirely inside one piece does not
                moved = (nnz - carved) + 2 * pieces + BS * NB * 4
                print(f"BS={BS:5d} NB={NB:4d} T={T}: carved {100 * carved / nnz:5.1f}% of edges in {pieces / 1e6:6.2f} M pieces "
                      f"-> L2->SM rows {100 * moved / nnz:5.1f}% of today")


if __name__ == "__main__":
    main()
