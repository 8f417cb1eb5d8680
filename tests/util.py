"""Shared test helpers: golden fixtures, the oracle, and a pure-numpy walker of the edge stream."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ppnp_oracle as oracle  # noqa: E402  (TEST INFRASTRUCTURE)


def load_std(name):
    z = np.load(os.path.join(GOLDEN, f"{name}_std.npz"))
    n = len(z["adj_indptr"]) - 1
    adj = sp.csr_matrix((np.ones(len(z["adj_indices"]), dtype=np.float32), z["adj_indices"], z["adj_indptr"]), shape=(n, n))
    return z, adj


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"{name}_golden.npz"))


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def epi_coef(epi, alpha, deg):
    oma = 1.0 - alpha
    if epi == 0:
        return oma, alpha
    if epi == 1:
        return oma / np.sqrt(deg), alpha / np.sqrt(deg)
    if epi == 2:
        return oma / deg, alpha / np.sqrt(deg)
    if epi == 3:
        return oma / np.sqrt(deg), alpha
    if epi == 4:
        return oma / deg, alpha
    raise ValueError(epi)


def walk_stream(plan, Zin, T, alpha, epi, use_vals, rows=None):
    """What csrc/appnp_spmm.cu computes from the plan arrays, edge by edge, in numpy fp64.
    Used to test the HOST-side plan logic without a GPU."""
    cols = plan.cols.cpu().numpy()
    vals = plan.vals.cpu().numpy() if plan.vals is not None else None
    seg_row = plan.seg_row.cpu().numpy()
    chunk_seg = plan.chunk_seg.cpu().numpy()
    W = plan.chunk_edges
    if plan.lane_group:   # lane-transposed storage (include/ppnp_b200.h): read logical position p from `stored`
        G = plan.lane_group
        SR = max(1, 16 // G)
        SE, CPS = SR * G, 4 // SR
        p = np.arange(W)
        j, r, l = p // SE, (p % SE) // G, p % G
        stored = (j // CPS) * (CPS * SE) + l * 4 + (j % CPS) * SR + r
        cols = cols.reshape(-1, W)[:, stored].reshape(-1)
        if vals is not None:
            vals = vals.reshape(-1, W)[:, stored].reshape(-1)
    F = Zin.shape[1]
    out = np.full((plan.n, F), np.nan)
    partial = np.full((max(plan.n_slots, 1), F), np.nan)
    written = np.zeros(plan.n, dtype=np.int32)
    for c in range(plan.n_chunks):
        s = int(chunk_seg[c])
        acc = np.zeros(F)
        cnt = 0
        for e in range(c * W, (c + 1) * W):
            raw = int(cols[e])
            col = raw & 0x7FFFFFFF
            acc += (vals[e] if use_vals else 1.0) * Zin[col]
            cnt += 1
            if raw < 0:
                sv = int(seg_row[s])
                s += 1
                if sv < 0:
                    partial[sv & 0x7FFFFFFF] = acc
                else:
                    a, b = epi_coef(epi, alpha, cnt)
                    out[sv] = a * acc + b * T[sv]
                    written[sv] += 1
                acc = np.zeros(F)
                cnt = 0
    fp = plan.fix_ptr.cpu().numpy()
    fr = plan.fix_row.cpu().numpy()
    fd = plan.fix_deg.cpu().numpy()
    for q in range(plan.n_fix):
        acc = partial[fp[q]:fp[q + 1]].sum(0)
        a, b = epi_coef(epi, alpha, float(fd[q]))
        out[fr[q]] = a * acc + b * T[fr[q]]
        written[fr[q]] += 1
    if rows is None:
        assert (written == 1).all(), "every row must be produced exactly once"
    else:      # a stream over a subset of the rows (plan.build_stream_plan(subset=True))
        assert (written[rows] == 1).all() and written.sum() == len(rows), "the listed rows exactly once, nothing else"
    return out


def walk_tiled(tp, Zin, T, alpha, epi, use_vals):
    """What csrc/appnp_tiled.cu computes from a TiledPlan, edge by edge, in numpy fp64 -- and the invariants the
    kernel relies on (piece counters, windows, no slot ending twice inside a slab, slots owned by one warp).
    Returns the rows the plan produces (NaN elsewhere)."""
    cols = tp.cols.cpu().numpy()
    vals = tp.vals.cpu().numpy() if tp.vals is not None else None
    meta = tp.slab_meta.cpu().numpy()
    pslot = tp.piece_slot.cpu().numpy()
    wptr = tp.warp_slab_ptr.cpu().numpy()
    cptr = tp.cta_slot_ptr.cpu().numpy()
    srow = tp.slot_row.cpu().numpy()
    rdeg = tp.row_deg.cpu().numpy()
    NW = tp.warps_per_cta
    F = Zin.shape[1]
    out = np.full((tp.n, F), np.nan)
    written = np.zeros(tp.n, dtype=np.int32)
    for cta in range(tp.n_ctas):
        ns = int(cptr[cta + 1] - cptr[cta])
        assert ns + 1 <= tp.slots_cap
        acc_s = np.zeros((ns + 1, F))
        owner = np.full(ns + 1, -1)
        for w in range(NW):
            s0, s1 = int(wptr[cta * NW + w]), int(wptr[cta * NW + w + 1])
            if s0 == s1:
                continue
            p = int(meta[s0, 0])
            acc = np.zeros(F)
            win = 0
            for s in range(s0, s1):
                assert int(meta[s, 0]) == p, "slab_meta piece counter out of step"
                assert (int(meta[s, 1]) & 0x3FFFFFFF) >= win, "windows must not decrease along a warp"
                win = int(meta[s, 1]) & 0x3FFFFFFF
                hazard = (int(meta[s, 1]) >> 30) & 1
                ended = []
                for e in range(s * 32, s * 32 + 32):
                    raw = int(cols[e])
                    acc += (vals[e] if use_vals else 1.0) * Zin[raw & 0x7FFFFFFF]
                    if raw < 0:
                        sl = int(pslot[p]); p += 1
                        assert 0 <= sl <= ns
                        if sl < ns:
                            assert owner[sl] in (-1, w), "a slot must belong to one warp"
                            owner[sl] = w
                            ended.append(sl)
                        acc_s[sl] += acc
                        acc = np.zeros(F)
                assert hazard == int(len(ended) != len(set(ended))), "slabs in which a slot ends twice must be marked (and only those)"
        sl = 0
        while sl < ns:
            row = int(srow[cptr[cta] + sl])
            assert row >= 0
            a = acc_s[sl].copy()
            t = sl + 1
            while t < ns and int(srow[cptr[cta] + t]) < 0:
                assert (int(srow[cptr[cta] + t]) & 0x7FFFFFFF) == row
                a += acc_s[t]; t += 1
            ca, cb = epi_coef(epi, alpha, float(rdeg[row]))
            out[row] = ca * a + cb * T[row]
            written[row] += 1
            sl = t
    hub = tp.hub_rows.cpu().numpy()
    assert (written[hub] == 1).all() and written.sum() == len(hub), "every hub row exactly once, nothing else"
    return out
