"""Plan of the shared-memory-resident hub-row step (csrc/appnp_tiled.cu, include/ppnp_b200.h section 2b).

The rows of the highest degrees ("hub rows") are dealt to CTAs; every CTA keeps one accumulator slot per row
(or per part of a very long row) in shared memory and its warps walk their own edge streams column window by
column window, so that a row of Z gathered by one warp is an L1 hit for the other warps of the SM while the
window is current.  Everything else (the rows of low degree) stays with the row-major stream
(``plan.build_stream_plan`` over the remaining rows).  Building the plan is index bookkeeping on the normalised
CSR -- sorts, prefix sums and scatters with torch tensor ops on the device the CSR lives on.

Reference anchor: like ``plan.py`` this is a re-encoding of the CSR that helpers.py:58-63 (calc_A_hat)
produces; which (row, column, value) triples exist never changes.
"""
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib
from .plan import FLAG_I32, StreamPlan, build_stream_plan, degree_order


@dataclass
class TiledPlan:
    n: int
    n_slabs: int
    n_pieces: int
    n_ctas: int
    warps_per_cta: int
    slots_cap: int               # max slots of a CTA + 1 (spare slot for padding pieces)
    slack: int
    cols: torch.Tensor           # int32 [n_slabs * 32]
    vals: Optional[torch.Tensor]
    slab_meta: torch.Tensor      # int32 [n_slabs, 2]: first piece index, window (numbered from 1)
    piece_slot: torch.Tensor     # int32 [n_pieces + 32]
    warp_slab_ptr: torch.Tensor  # int32 [n_ctas * warps + 1]
    cta_slot_ptr: torch.Tensor   # int32 [n_ctas + 1]
    slot_row: torch.Tensor       # int32 [total slots]
    row_deg: torch.Tensor        # fp32 [n]
    hub_rows: torch.Tensor       # int64: rows this plan produces
    rest: Optional[StreamPlan]   # row-major stream over all other rows (None when every row is a hub row)
    window_ends: torch.Tensor    # int64: column-rank boundaries of the windows
    stats: dict = field(default_factory=dict)
    _struct: object = field(default=None, repr=False)

    @property
    def device(self):
        return self.cols.device

    def struct(self):
        if self._struct is None:
            s = _lib.TiledPlanStruct()
            s.n, s.n_slabs, s.n_pieces = self.n, self.n_slabs, self.n_pieces
            s.n_ctas, s.warps_per_cta, s.slots_cap, s.slack = self.n_ctas, self.warps_per_cta, self.slots_cap, self.slack
            s.cols = self.cols.data_ptr()
            s.vals = self.vals.data_ptr() if self.vals is not None else None
            s.slab_meta = self.slab_meta.data_ptr()
            s.piece_slot = self.piece_slot.data_ptr()
            s.warp_slab_ptr = self.warp_slab_ptr.data_ptr()
            s.cta_slot_ptr = self.cta_slot_ptr.data_ptr()
            s.slot_row = self.slot_row.data_ptr()
            s.row_deg = self.row_deg.data_ptr()
            self._struct = s
        return self._struct

    def index_bytes(self):
        b = self.cols.numel() * 4 + self.slab_meta.numel() * 4 + self.piece_slot.numel() * 4
        if self.rest is not None:
            b += self.rest.index_bytes()
        return b


def _lpt(sizes, counts, n_bins, cap, group=None):
    """Longest-processing-time-first: items (descending ``sizes``, each taking ``counts[i]`` of a bin's ``cap``
    places) go to the bin with the least load that still has room.  With ``group`` (non-decreasing bin-group id
    per item, bins numbered inside their group) every group is balanced on its own.  Host-side loop over a few
    10^4 .. 10^5 items (hub rows, accumulator slots); returns the bin of every item."""
    import heapq
    sz = sizes.tolist()
    ct = counts.tolist() if counts is not None else None
    gr = group.tolist() if group is not None else None
    out = [0] * len(sz)
    heaps = {}
    for i, s_ in enumerate(sz):
        g_ = gr[i] if gr is not None else 0
        h = heaps.get(g_)
        if h is None:
            h = [(0, b, 0) for b in range(n_bins)]
            heaps[g_] = h
        c_ = ct[i] if ct is not None else 1
        skipped = []
        while True:
            if not h:
                raise ValueError("no bin has room left")
            load, b, used = heapq.heappop(h)
            if used + c_ <= cap:
                break
            skipped.append((load, b, used))
        out[i] = b
        heapq.heappush(h, (load + s_, b, used + c_))
        for x in skipped:
            heapq.heappush(h, x)
    return out


def choose_windows(crank_hist, bucket, n_ctas, fine_cols, fine_min_reuse, coarse_edges, n_cols):
    """Column-rank boundaries of the windows.  ``crank_hist[b]`` = hub edges whose column rank lies in bucket b
    (``bucket`` ranks wide).  Fine windows of ``fine_cols`` ranks run from rank 0 while a CTA's share of a
    window holds at least ``fine_min_reuse`` edges per column (the zone where the L1 pays); after that the
    windows are cut at equal numbers of edges (``coarse_edges`` per CTA): they only keep the warps of a CTA and
    the CTAs of the chip inside one region of Z at a time (L2 locality of the cold gathers)."""
    dev = crank_hist.device
    c = torch.cumsum(crank_hist.to(torch.float64), 0) / float(n_ctas)          # edges per CTA up to the end of bucket b
    c = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), c])
    per = max(1, fine_cols // bucket)
    nb = int(crank_hist.numel())
    starts = torch.arange(0, nb, per, device=dev)
    stops = torch.clamp(starts + per, max=nb)
    e_win = c[stops] - c[starts]
    ok = e_win >= fine_min_reuse * fine_cols
    # fine zone = the leading run of windows that pay
    bad = torch.nonzero(~ok).flatten()
    n_fine = int(bad[0].item()) if bad.numel() else int(starts.numel())
    ends = [int(min((i + 1) * per * bucket, n_cols)) for i in range(n_fine)]
    lo_b = min(n_fine * per, nb)
    if lo_b < nb:
        rest_edges = float(c[nb] - c[lo_b])
        k = max(1, int(round(rest_edges / max(coarse_edges, 1))))
        targets = c[lo_b] + (torch.arange(1, k + 1, device=dev, dtype=torch.float64) * (rest_edges / k))
        cut = torch.searchsorted(c[1:].contiguous(), targets.contiguous(), right=False) + 1     # bucket count
        cut = torch.clamp(cut, min=lo_b + 1, max=nb)
        cut = torch.unique(cut)
        ends += [int(min(int(x) * bucket, n_cols)) for x in cut.tolist()]
    if not ends or ends[-1] < n_cols:
        ends.append(n_cols)
    ends = sorted(set(ends))
    return torch.tensor(ends, dtype=torch.int64, device=dev), n_fine


def build_tiled_plan(indptr, indices, vals=None, *, n_ctas, warps_per_cta=16, slot_rows=395, min_hub_degree=32, part_div=4,
                     fine_cols=256, fine_min_reuse=1.5, coarse_edges=32768, slack=1, rest_chunk_edges=256, max_hub_rows=None,
                     build_rest=True):
    """Split the rows into hub rows (tiled plan) and the rest (row-major stream).

    n_ctas        row groups = CTAs along grid.x (the SMs of the device divided by the number of feature slices)
    slot_rows     accumulator slots of a CTA (shared-memory budget / slice bytes), the spare slot not counted
    min_hub_degree rows below this degree are never hub rows (nothing to re-use)
    part_div      a row longer than (hub edges per warp) / part_div is cut into parts, one slot each
    fine_cols, fine_min_reuse, coarse_edges: see ``choose_windows``
    slack         windows a warp may run ahead of the slowest warp of its CTA
    """
    dev = indices.device
    n = int(indptr.numel()) - 1
    ip = indptr.to(torch.int64)
    nnz = int(ip[-1].item())
    if nnz != int(indices.numel()):
        raise ValueError("indptr[-1] != len(indices)")
    if not (1 <= warps_per_cta <= 16):
        raise ValueError("1..16 warps per CTA")
    NW = int(warps_per_cta)
    deg = ip[1:] - ip[:-1]
    if bool((deg <= 0).any()):
        raise ValueError("every row needs at least one edge (A_hat rows hold their self loop)")
    order = degree_order(indptr)
    rank = torch.empty(n, dtype=torch.int64, device=dev)
    rank[order] = torch.arange(n, device=dev, dtype=torch.int64)
    sdeg = deg[order]

    # ---- hub rows: as many of the highest degrees as the slots hold
    n_cand = int((sdeg >= min_hub_degree).sum().item())
    if max_hub_rows is not None:
        n_cand = min(n_cand, int(max_hub_rows))
    if n_cand == 0:
        raise ValueError("no row reaches min_hub_degree: nothing to tile")
    shrink = 1.0
    while True:
        Nh = n_cand
        total = int(sdeg[:Nh].sum().item())
        part_max = max(32, int(total / (n_ctas * NW) / part_div))
        nparts = torch.div(sdeg[:Nh] + part_max - 1, part_max, rounding_mode="floor")
        cap_total = int(n_ctas * slot_rows * shrink)
        cs = torch.cumsum(nparts, 0)
        if int(cs[-1].item()) > cap_total:
            Nh = int(torch.searchsorted(cs, torch.tensor([cap_total], device=dev, dtype=cs.dtype), right=True).item())
            if Nh == 0:
                raise ValueError("slot budget too small for a single hub row")
            total = int(sdeg[:Nh].sum().item())
            part_max = max(32, int(total / (n_ctas * NW) / part_div))
            nparts = torch.div(sdeg[:Nh] + part_max - 1, part_max, rounding_mode="floor")
        pos = torch.arange(Nh, device=dev, dtype=torch.int64)
        try:
            cta_of = torch.tensor(_lpt(sdeg[:Nh].cpu(), nparts.cpu(), n_ctas, slot_rows), dtype=torch.int64, device=dev)
        except ValueError:
            shrink *= 0.97
            n_cand = Nh
            continue
        slots_per_cta = torch.zeros(n_ctas, dtype=torch.int64, device=dev).index_add_(0, cta_of, nparts)
        break
    hub_rows = order[:Nh]
    hd = sdeg[:Nh]

    # ---- slots: rows of a CTA in rank order, the parts of a row adjacent
    key = cta_of * Nh + pos
    perm = torch.argsort(key)
    np_sorted = nparts[perm]
    first_slot_sorted = torch.cumsum(np_sorted, 0) - np_sorted             # global slot index of a row's first part
    gslot_base = torch.empty(Nh, dtype=torch.int64, device=dev)
    gslot_base[perm] = first_slot_sorted
    cta_slot_ptr = torch.zeros(n_ctas + 1, dtype=torch.int64, device=dev)
    cta_slot_ptr[1:] = torch.cumsum(slots_per_cta, 0)
    n_slots_total = int(cta_slot_ptr[-1].item())
    slot_owner = torch.repeat_interleave(perm, np_sorted)                  # hub position of every global slot
    slot_part = torch.arange(n_slots_total, device=dev, dtype=torch.int64) - first_slot_sorted.repeat_interleave(np_sorted)
    slot_row = hub_rows[slot_owner].to(torch.int32)
    slot_row = torch.where(slot_part > 0, slot_row | FLAG_I32, slot_row).to(torch.int32)

    # ---- hub edges, sorted inside every row by column rank; the part of every edge
    L = hd
    a = torch.cumsum(L, 0) - L
    E = int(L.sum().item())
    e_hub = torch.repeat_interleave(pos, L)
    src = torch.repeat_interleave(ip[:-1][hub_rows] - a, L) + torch.arange(E, device=dev, dtype=torch.int64)
    e_col = indices[src].to(torch.int64)
    e_val = None if vals is None else vals[src].to(torch.float32)
    e_crank = rank[e_col]
    k1 = e_hub * n + e_crank
    k1, p1 = torch.sort(k1)
    e_hub, e_col, e_crank = e_hub[p1], e_col[p1], e_crank[p1]
    if e_val is not None:
        e_val = e_val[p1]
    del src, p1, k1
    pos_in_row = torch.arange(E, device=dev, dtype=torch.int64) - a[e_hub]
    e_part = torch.div(pos_in_row * nparts[e_hub], L[e_hub], rounding_mode="floor")
    e_gslot = gslot_base[e_hub] + e_part
    del pos_in_row, e_part
    slot_edges = torch.bincount(e_gslot, minlength=n_slots_total)

    # ---- slots to warps: inside a CTA by descending size, each to the warp with the least edges so far
    slot_cta = torch.repeat_interleave(torch.arange(n_ctas, device=dev, dtype=torch.int64), slots_per_cta)
    big = int(slot_edges.max().item()) + 1
    sk = slot_cta * big + (big - 1 - slot_edges)
    sp = torch.argsort(sk, stable=True)
    warp_of_slot = torch.empty(n_slots_total, dtype=torch.int64, device=dev)
    warp_of_slot[sp] = torch.tensor(_lpt(slot_edges[sp].cpu(), None, NW, 1 << 30, group=slot_cta[sp].cpu()), dtype=torch.int64, device=dev)
    slot_local = torch.arange(n_slots_total, device=dev, dtype=torch.int64) - cta_slot_ptr[:-1][slot_cta]

    # ---- windows
    bucket = 32
    hist = torch.bincount(torch.div(e_crank, bucket, rounding_mode="floor"), minlength=(n + bucket - 1) // bucket)
    window_ends, n_fine = choose_windows(hist, bucket, n_ctas, fine_cols, fine_min_reuse, coarse_edges, n)
    n_win = int(window_ends.numel())
    e_win = torch.bucketize(e_crank, window_ends, right=True)
    if int(e_win.max().item()) >= n_win:
        raise AssertionError("window boundaries do not cover the column space")

    # ---- stream order: (warp, window, slot, column rank)
    e_wid = slot_cta[e_gslot] * NW + warp_of_slot[e_gslot]
    e_sl = slot_local[e_gslot]
    cap_bits = max(1, int(slot_rows + 1).bit_length())
    n_bits = max(1, int(n).bit_length())
    w_bits = max(1, int(n_win).bit_length())
    if (int(n_ctas * NW).bit_length() + w_bits + cap_bits + n_bits) > 62:
        raise ValueError("sort key does not fit 62 bits")
    k2 = (((e_wid << w_bits) | e_win) << cap_bits | e_sl) << n_bits | e_crank
    k2, p2 = torch.sort(k2)
    e_wid, e_win, e_sl, e_col = e_wid[p2], e_win[p2], e_sl[p2], e_col[p2]
    if e_val is not None:
        e_val = e_val[p2]
    piece_key = k2 >> n_bits
    del k2, p2, e_crank, e_gslot, e_hub
    last = torch.ones(E, dtype=torch.bool, device=dev)
    last[:-1] = piece_key[1:] != piece_key[:-1]
    first = torch.ones(E, dtype=torch.bool, device=dev)
    first[1:] = last[:-1]
    piece_of = torch.cumsum(first.to(torch.int64), 0) - 1                  # piece index of every edge
    n_real_pieces = int(piece_of[-1].item()) + 1
    piece_first_edge = torch.nonzero(first).flatten()
    piece_wid = e_wid[piece_first_edge]
    del piece_key

    n_warps = n_ctas * NW
    warp_edges = torch.bincount(e_wid, minlength=n_warps)
    warp_first_edge = torch.cumsum(warp_edges, 0) - warp_edges
    # ---- layout: every warp starts on a slab boundary.  A slab in which one slot ends twice (tiny pieces of one
    # row in consecutive windows) is marked: the kernel then adds the piece ends of that slab one lane group at
    # a time instead of all in one instruction.
    warp_len = ((warp_edges + 31) // 32) * 32
    warp_off = torch.cumsum(warp_len, 0) - warp_len
    e_pos = warp_off[e_wid] + (torch.arange(E, device=dev, dtype=torch.int64) - warp_first_edge[e_wid])
    end_pos = e_pos[last]
    end_slot = e_sl[last]
    total_len = int((warp_off[-1] + warp_len[-1]).item())
    n_slabs = total_len // 32
    hk = (torch.div(end_pos, 32, rounding_mode="floor") << cap_bits) | end_slot
    hs = torch.sort(hk).values
    dup = torch.zeros_like(hs, dtype=torch.bool)
    dup[1:] = hs[1:] == hs[:-1]
    hazard = torch.zeros(n_slabs, dtype=torch.bool, device=dev)
    hazard[hs[dup] >> cap_bits] = True
    cols = torch.zeros(total_len, dtype=torch.int32, device=dev)
    cols[e_pos] = e_col.to(torch.int32)
    cols[end_pos] |= FLAG_I32
    svals = None
    if e_val is not None:
        svals = torch.zeros(total_len, dtype=torch.float32, device=dev)
        svals[e_pos] = e_val
    win_pos = torch.zeros(total_len, dtype=torch.int32, device=dev)
    win_pos[e_pos] = (e_win + 1).to(torch.int32)
    piece_slot = end_slot.to(torch.int32)                                   # stream order == piece order
    n_pieces = n_real_pieces
    piece_slot = torch.cat([piece_slot, torch.zeros(32, dtype=torch.int32, device=dev)])
    ends_per_slab = (cols < 0).view(n_slabs, 32).sum(1)
    piece0 = torch.cumsum(ends_per_slab, 0) - ends_per_slab
    slab_win = win_pos.view(n_slabs, 32).max(1).values.to(torch.int64)
    slab_meta = torch.stack([piece0.to(torch.int32), (slab_win | (hazard.to(torch.int64) << 30)).to(torch.int32)], 1).contiguous()
    warp_slab_ptr = torch.zeros(n_warps + 1, dtype=torch.int64, device=dev)
    warp_slab_ptr[:-1] = torch.div(warp_off, 32, rounding_mode="floor")
    warp_slab_ptr[-1] = n_slabs
    if n_slabs >= (1 << 31) or n_pieces >= (1 << 31):
        raise ValueError("tiled stream too long for 32-bit slab / piece indices")

    rest = None
    if build_rest and Nh < n:
        rest = build_stream_plan(indptr, indices, vals, rest_chunk_edges, order=order[Nh:], subset=True)
    cta_edges = torch.zeros(n_ctas, dtype=torch.int64, device=dev).index_add_(0, torch.div(e_wid, NW, rounding_mode="floor"),
                                                                           torch.ones_like(e_wid))
    stats = {
        "hub_rows": Nh, "hub_edges": E, "hub_edge_share": E / max(nnz, 1), "min_hub_degree_taken": int(hd[-1].item()),
        "slots": n_slots_total, "split_rows": int((nparts > 1).sum().item()), "part_max": part_max,
        "pieces": n_real_pieces, "edges_per_piece": E / max(n_real_pieces, 1), "hazard_slabs": int(hazard.sum().item()),
        "stream_edges": total_len, "padding_share": 1.0 - E / max(total_len, 1),
        "windows": n_win, "fine_windows": n_fine, "fine_cols": fine_cols,
        "cta_edges_max_over_mean": float(cta_edges.max().item()) / max(float(cta_edges.double().mean().item()), 1.0),
        "warp_slabs_max_over_mean": float(warp_len.max().item()) / max(float(warp_len.double().mean().item()), 1.0),
        "slots_per_cta_max": int(slots_per_cta.max().item()),
    }
    return TiledPlan(n=n, n_slabs=n_slabs, n_pieces=n_pieces, n_ctas=int(n_ctas), warps_per_cta=NW, slots_cap=int(slot_rows) + 1,
                     slack=int(slack), cols=cols, vals=svals, slab_meta=slab_meta, piece_slot=piece_slot,
                     warp_slab_ptr=warp_slab_ptr.to(torch.int32), cta_slot_ptr=cta_slot_ptr.to(torch.int32), slot_row=slot_row,
                     row_deg=deg.to(torch.float32), hub_rows=hub_rows, rest=rest, window_ends=window_ends, stats=stats)
