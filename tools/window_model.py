"""DESIGN TOOL (not product code): what does re-ordering WHOLE CHUNKS of the hub rows buy?

A hub row is already cut into 256-edge chunks, each a partial segment with its own slot, so its chunks can be
processed in any order at no extra partial-sum traffic.  Candidate: sort every row's columns by degree rank, then
process the single-segment chunks sorted by the rank of their first column ("window order"), so that the chunks an
SM (and the chip) works on at one time cover the same hot columns.  This script feeds the column streams to the LRU
model tools/l1sim.c (per-SM L1 in front of a shared L2), like tools/carve_model.py.

  python tools/window_model.py [l1_rows] [l2_rows]
"""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ppnp_oracle as oracle  # noqa: E402
from ppnp_b200.plan import build_stream_plan, degree_order, window_order_chunks, rank_sorted_csr  # noqa: E402

SIM = os.path.join(ROOT, "tools", "_build", "l1sim")


def misses(plan, cache_rows, sms, unit, l2_rows):
    tmp = "/tmp/window_model_cols.i32"
    plan.cols.numpy().tofile(tmp)
    out = subprocess.run([SIM, tmp, str(plan.n_chunks), str(plan.chunk_edges), str(plan.n), str(cache_rows), str(sms), str(unit), str(l2_rows)],
                         capture_output=True, text=True, check=True).stdout
    return int(out.split("L1 misses")[1].split("(")[0]), int(out.split("L2 misses")[1].split("(")[0])


def main():
    l1_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    l2_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 190_000
    subprocess.run(["make", "-C", os.path.join(ROOT, "tools"), "_build/l1sim"], check=True, capture_output=True)
    ip, idx = oracle.rmat_graph(2_000_000, 26_400_000, 21)
    oip, oidx, _, _ = oracle.c_a_hat(ip, idx, None, "sym")
    tip, tidx = torch.from_numpy(oip.astype(np.int32)), torch.from_numpy(oidx)
    order = degree_order(tip)
    base = build_stream_plan(tip, tidx, None, 256, order)
    sidx, _, crank = rank_sorted_csr(tip, tidx, None)
    ranked = build_stream_plan(tip, sidx, None, 256, order)
    cases = [("degree order (bench default)", base), ("+ columns of a row sorted by rank", ranked)]
    for key in ("first", "mid", "last"):
        cases.append((f"+ single-segment chunks in window order ({key} column)", window_order_chunks(ranked, crank, key=key)))
    for unit in (64, 16):
        for tag, p in cases:
            m1, m2 = misses(p, l1_rows, 148, unit, l2_rows)
            print(f"unit {unit:3d} | {tag}: {100 * m1 / p.nnz:.1f}% of the edges cross L2 -> SM, {m2 / 1e6:.2f} M rows come from HBM "
                  f"({p.n_slots / 1e6:.2f} M partial rows)", flush=True)


if __name__ == "__main__":
    main()
