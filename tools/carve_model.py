"""DESIGN TOOL (not product code): rank carved-stream parameters on the CPU before spending GPU time.

Builds the config-4 R-MAT graph with the oracle's generator, the edge streams with ppnp_b200/plan.py (torch
on the CPU) and feeds their column stream to the LRU model tools/l1sim.c:

  python tools/carve_model.py l1     per-SM L1 (148 SMs, units of 64 chunks): rows crossing L2 -> SM
  python tools/carve_model.py l2     one shared cache: rows crossing HBM -> L2, for 256-byte rows (config 4,
                                     96 MB) and for the config-5 scale model (64-byte rows, cache / 8)
  python tools/carve_model.py l12    per-SM L1 (512 rows) in front of a shared L2 (375 k rows): one- and
                                     two-level carves (L1-sized blocks over the hottest columns, L2-sized
                                     blocks over the rest), with and without chunk interleaving

Round-1 output is quoted in profiles/r01_variants.md and DESIGN.md section 8.
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
from ppnp_b200.plan import build_carved_plan, build_stream_plan, degree_order  # noqa: E402

SIM = os.path.join(ROOT, "tools", "_build", "l1sim")


def misses(plan, cache_rows, sms, unit, l2_rows=0):
    tmp = "/tmp/carve_model_cols.i32"
    plan.cols.numpy().tofile(tmp)
    out = subprocess.run([SIM, tmp, str(plan.n_chunks), str(plan.chunk_edges), str(plan.n), str(cache_rows), str(sms), str(unit)] +
                         ([str(l2_rows)] if l2_rows else []), capture_output=True, text=True, check=True).stdout
    m1 = int(out.split("L1 misses")[1].split("(")[0])
    return (m1, int(out.split("L2 misses")[1].split("(")[0])) if l2_rows else m1


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "l1"
    subprocess.run(["make", "-C", os.path.join(ROOT, "tools"), "_build/l1sim"], check=True, capture_output=True)
    ip, idx = oracle.rmat_graph(2_000_000, 26_400_000, 21)
    oip, oidx, _, _ = oracle.c_a_hat(ip, idx, None, "sym")
    tip, tidx = torch.from_numpy(oip.astype(np.int32)), torch.from_numpy(oidx)
    base = build_stream_plan(tip, tidx, None, 256, degree_order(tip))
    if mode == "l1":
        for rows in (256, 512, 768):
            print(f"degree order, L1 {rows} rows: {100 * misses(base, rows, 148, 64) / base.nnz:.1f}% of the edges cross L2 -> SM")
        for bc, nb, t in ((512, 64, 4), (512, 64, 8), (512, 64, 16), (384, 96, 4), (512, 256, 4), (768, 64, 16)):
            p = build_carved_plan(tip, tidx, None, 256, bc, nb, t)
            for rows in (512, 700):
                print(f"carve {bc}x{nb} min piece {t}, L1 {rows} rows: carved {100 * p.carve['carved_edges'] / p.nnz:.1f}% in "
                      f"{p.carve['carved_pieces'] / 1e6:.2f} M pieces, {100 * misses(p, rows, 148, 64) / p.nnz:.1f}% of the edges cross "
                      f"L2 -> SM (+{100 * p.n_slots / p.nnz:.1f}% partial reads)", flush=True)
    elif mode == "l12":
        cases = [("degree order", base),
                 ("carve 512x64 min 8", build_carved_plan(tip, tidx, None, 256, 512, 64, 8)),
                 ("carve 512x64 min 8, interleaved", build_carved_plan(tip, tidx, None, 256, 512, 64, 8, interleave=True)),
                 ("carve 512x64 min 8 + 125k x 16 min 16", build_carved_plan(tip, tidx, None, 256, levels=[(512, 64, 8), (125_000, 16, 16)])),
                 ("carve 512x128 min 6 + 250k x 8 min 16", build_carved_plan(tip, tidx, None, 256, levels=[(512, 128, 6), (250_000, 8, 16)]))]
        for tag, p in cases:
            m1, m2 = misses(p, 512, 148, 64, 375_000)
            print(f"{tag}: {100 * m1 / p.nnz:.1f}% of the edges cross L2 -> SM, {m2 / 1e6:.2f} M rows come from HBM, "
                  f"{p.n_slots / 1e6:.2f} M partial rows", flush=True)
    else:
        for cache_rows, tag in ((190_000, "config-5 scale model: 64-byte rows, cache / 8"), (375_000, "config 4: 256-byte rows, 96 MB")):
            m = misses(base, cache_rows, 1, 64)
            print(f"{tag}\n  degree order: {m / 1e6:.2f} M gathered rows miss + {base.n_slots / 1e6:.2f} M partial rows = "
                  f"{(m + 2 * base.n_slots) / base.n:.2f} x n rows through HBM")
            for bc, nb, t in ((125_000, 16, 8), (125_000, 16, 16), (125_000, 16, 32), (62_500, 32, 16), (250_000, 8, 16)):
                p = build_carved_plan(tip, tidx, None, 256, bc, nb, t, wide_cta=False)
                m = misses(p, cache_rows, 1, 64)
                print(f"  carve {bc}x{nb} min piece {t}: carved {100 * p.carve['carved_edges'] / p.nnz:.0f}%, {m / 1e6:.2f} M misses + "
                      f"{p.n_slots / 1e6:.2f} M partial rows = {(m + 2 * p.n_slots) / p.n:.2f} x n rows through HBM", flush=True)


if __name__ == "__main__":
    main()
