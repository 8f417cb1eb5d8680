"""Edge-stream plan for the propagation kernel (csrc/appnp_spmm.cu).

The kernel balances load by giving every lane group the same number of EDGES, not rows: the rows
of A_hat are laid end to end in processing order ("the stream") and cut into chunks of
``chunk_edges`` edges; a row that crosses a chunk boundary is cut there and its pieces become
partial segments whose sums a fix-up kernel adds in order.  Building the stream is index
bookkeeping over the normalised CSR (the output of ``ppnp_csr_normalize``): prefix sums and
gathers, done here with torch tensor ops on whatever device the CSR lives on (PyTorch is the
plumbing; nothing below touches feature data).

Reference anchor: the stream is a re-encoding of the CSR that helpers.py:58-63 (calc_A_hat)
produces; processing order and cuts never change which (row, column, value) triples exist.
"""
from dataclasses import dataclass, field
from typing import Optional

import os

import torch

from . import _lib

FLAG_I32 = -(1 << 31)          # PPNP_FLAG as a signed int32
PLAN_WIDE_CTA = 1              # ppnp_plan_t.flags bit 0 (include/ppnp_b200.h)
PLAN_LANE_GROUP_SHIFT = 8      # ppnp_plan_t.flags bits 8..15: lane group of a lane-transposed stream


@dataclass
class StreamPlan:
    n: int
    nnz: int                    # real edges in the stream (self loops included)
    chunk_edges: int
    n_chunks: int
    cols: torch.Tensor          # int32 [n_chunks * chunk_edges]
    vals: Optional[torch.Tensor]  # fp32 same length, or None
    seg_row: torch.Tensor       # int32 [n_segs]
    chunk_seg: torch.Tensor     # int32 [n_chunks]
    fix_ptr: torch.Tensor       # int32 [n_fix + 1]
    fix_row: torch.Tensor       # int32 [n_fix]
    fix_deg: torch.Tensor       # fp32 [n_fix]
    n_slots: int
    order: Optional[torch.Tensor] = None   # processing order of the rows (None = natural)
    n_segs_real: int = 0
    row_deg: Optional[torch.Tensor] = None  # fp32 [n]: full row degrees for partial-row streams
    wide_cta: bool = False      # PPNP_PLAN_WIDE_CTA: one 1024-thread CTA per SM walks consecutive chunks
    lane_group: int = 0         # > 0: cols/vals are stored lane-transposed for groups of this many lanes
    carve: Optional[dict] = None  # statistics of build_carved_plan (None for row-major streams)
    _struct: object = field(default=None, repr=False)

    @property
    def n_segs(self):
        return self.n_segs_real if self.n_segs_real else int(self.seg_row.numel())

    @property
    def n_fix(self):
        return int(self.fix_row.numel())

    @property
    def device(self):
        return self.cols.device

    def struct(self):
        """ctypes mirror of ppnp_plan_t (cached; keeps the tensors alive through self)."""
        if self._struct is None:
            s = _lib.PlanStruct()
            s.n, s.n_edges, s.n_chunks = self.n, self.n_chunks * self.chunk_edges, self.n_chunks
            s.n_segs, s.n_fix, s.n_slots = self.n_segs, self.n_fix, self.n_slots
            s.chunk_edges = self.chunk_edges
            s.flags = (PLAN_WIDE_CTA if self.wide_cta else 0) | (self.lane_group << PLAN_LANE_GROUP_SHIFT)
            s.cols = self.cols.data_ptr()
            s.vals = self.vals.data_ptr() if self.vals is not None else None
            s.seg_row = self.seg_row.data_ptr()
            s.chunk_seg = self.chunk_seg.data_ptr()
            s.fix_ptr = self.fix_ptr.data_ptr()
            s.fix_row = self.fix_row.data_ptr() if self.n_fix else None
            s.fix_deg = self.fix_deg.data_ptr() if self.n_fix else None
            s.row_deg = self.row_deg.data_ptr() if self.row_deg is not None else None
            self._struct = s
        return self._struct

    def index_bytes(self):
        """Bytes of index data one step streams (cols + seg_row + chunk_seg [+ vals])."""
        b = self.cols.numel() * 4 + self.seg_row.numel() * 4 + self.chunk_seg.numel() * 4
        return b

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return StreamPlan(self.n, self.nnz, self.chunk_edges, self.n_chunks, mv(self.cols), mv(self.vals),
                          mv(self.seg_row), mv(self.chunk_seg), mv(self.fix_ptr), mv(self.fix_row),
                          mv(self.fix_deg), self.n_slots, mv(self.order), self.n_segs_real, mv(self.row_deg),
                          self.wide_cta, self.lane_group, self.carve)


def build_stream_plan_cuda(indptr, indices, vals=None, chunk_edges=256, order=None, subset=False, row_deg=None, lane_group=0):
    """``build_stream_plan`` (and, with ``lane_group`` > 0, ``lane_transpose`` on top of it) by the CUDA builder
    csrc/plan_build.cu: four scans over the listed rows and one pass over the edges instead of ~15 eager tensor
    passes over int64 temporaries of nnz entries.  Same arrays, bit for bit (tests/test_gpu_plan_build.py)."""
    if chunk_edges % 128 != 0 or chunk_edges <= 0:
        raise ValueError("chunk_edges must be a positive multiple of 128")
    if not indices.is_cuda:
        raise RuntimeError("build_stream_plan_cuda: CUDA tensors required")
    lib = _lib.load()
    dev = indices.device
    n = int(indptr.numel()) - 1
    ip = indptr.to(device=dev, dtype=torch.int64).contiguous()
    idx32 = indices.to(torch.int32).contiguous()
    v32 = None if vals is None else vals.to(torch.float32).contiguous()
    if order is not None:
        order = order.to(device=dev, dtype=torch.int64).contiguous()
        if not subset and order.numel() != n:
            raise ValueError("order must list every row once (pass subset=True for a partial stream)")
    m = n if order is None else int(order.numel())
    if m == 0:
        raise ValueError("no rows to stream")
    if row_deg is not None:
        row_deg = row_deg.to(device=dev, dtype=torch.float32).contiguous()
    W = int(chunk_edges)
    row_start = torch.empty(m + 1, dtype=torch.int64, device=dev)
    seg_first = torch.empty(m + 1, dtype=torch.int32, device=dev)
    slot_first = torch.empty(m + 1, dtype=torch.int32, device=dev)
    fix_first = torch.empty(m + 1, dtype=torch.int32, device=dev)
    totals = torch.empty(5, dtype=torch.int64, device=dev)
    ws_bytes = int(lib.ppnp_plan_workspace_bytes(m))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.ppnp_plan_measure(_lib.ptr(ip), _lib.ptr(order), m, W, _lib.ptr(row_start), _lib.ptr(seg_first), _lib.ptr(slot_first),
                                   _lib.ptr(fix_first), _lib.ptr(totals), _lib.ptr(ws), ws_bytes, _lib.current_stream())
        _lib.check(rc, "ppnp_plan_measure")
        nnz, n_segs, n_slots, n_fix, n_empty, nnz_csr = totals.tolist() + [int(ip[-1].item())]
        if nnz_csr != int(indices.numel()):
            raise ValueError("indptr[-1] != len(indices)")
        if n_empty:
            raise ValueError("every streamed row needs at least one edge (A_hat rows hold their self loop)")
        if nnz // W + m >= (1 << 31):
            raise ValueError("stream too long for 32-bit segment indices")
        n_chunks = ((((nnz + W - 1) // W) + 31) // 32) * 32
        cols = torch.empty(n_chunks * W, dtype=torch.int32, device=dev)
        svals = None if v32 is None else torch.empty(n_chunks * W, dtype=torch.float32, device=dev)
        # 64 spare entries: the kernel prefetches seg_row[s + lane] without a bounds check
        seg_row = torch.empty(n_segs + 64, dtype=torch.int32, device=dev)
        seg_row[n_segs:] = 0
        chunk_seg = torch.empty(n_chunks, dtype=torch.int32, device=dev)
        fix_ptr = torch.empty(n_fix + 1, dtype=torch.int32, device=dev)
        fix_row = torch.empty(n_fix, dtype=torch.int32, device=dev)
        fix_deg = torch.empty(n_fix, dtype=torch.float32, device=dev)
        rc = lib.ppnp_plan_fill(_lib.ptr(ip), _lib.ptr(idx32), _lib.ptr(v32), _lib.ptr(order), m, W, int(lane_group), _lib.ptr(row_start),
                                _lib.ptr(seg_first), _lib.ptr(slot_first), _lib.ptr(fix_first), _lib.ptr(row_deg), nnz, n_chunks,
                                n_segs, n_fix, _lib.ptr(cols), _lib.ptr(svals), _lib.ptr(seg_row), _lib.ptr(chunk_seg),
                                _lib.ptr(fix_ptr), _lib.ptr(fix_row) if n_fix else None, _lib.ptr(fix_deg) if n_fix else None,
                                _lib.current_stream())
        _lib.check(rc, "ppnp_plan_fill")
    return StreamPlan(n=n, nnz=nnz, chunk_edges=W, n_chunks=n_chunks, cols=cols, vals=svals, seg_row=seg_row, chunk_seg=chunk_seg,
                      fix_ptr=fix_ptr, fix_row=fix_row, fix_deg=fix_deg, n_slots=n_slots, order=order, n_segs_real=n_segs,
                      row_deg=row_deg, lane_group=int(lane_group))


def _cuda_builder_enabled():
    return os.environ.get("PPNP_PLAN_BUILDER", "cuda").lower() != "torch"


def build_stream_plan(indptr, indices, vals=None, chunk_edges=256, order=None, subset=False, row_deg=None):
    """Cut the CSR (indptr[n+1], indices[nnz], optional vals[nnz]) into the edge stream.

    ``order`` (int64, a permutation of the rows) is the processing order; column ids are never
    relabelled, so inputs and outputs of the propagation keep the caller's row order.  With
    ``subset=True`` ``order`` may list only some rows: the stream then produces exactly those rows
    (used to split a shard into interior and boundary rows, ppnp_b200/dist.py).

    CUDA tensors go through the CUDA builder (``build_stream_plan_cuda``; ``PPNP_PLAN_BUILDER=torch`` keeps the
    tensor-op form below, which is also the host mirror the CPU tests walk and the specification the CUDA
    builder is checked against bit for bit).
    """
    if indices.is_cuda and _cuda_builder_enabled():
        return build_stream_plan_cuda(indptr, indices, vals, chunk_edges, order, subset, row_deg)
    return build_stream_plan_torch(indptr, indices, vals, chunk_edges, order, subset, row_deg)


def build_stream_plan_torch(indptr, indices, vals=None, chunk_edges=256, order=None, subset=False, row_deg=None):
    """``build_stream_plan`` with torch tensor ops on whatever device the CSR lives on (host mirror / specification)."""
    if chunk_edges % 128 != 0 or chunk_edges <= 0:
        raise ValueError("chunk_edges must be a positive multiple of 128")
    dev = indices.device
    n = int(indptr.numel()) - 1
    ip = indptr.to(torch.int64)
    nnz = int(ip[-1].item())
    if nnz != int(indices.numel()):
        raise ValueError("indptr[-1] != len(indices)")
    W = chunk_edges
    deg = ip[1:] - ip[:-1]
    listed = deg if (order is None or not subset) else deg[order.to(device=dev, dtype=torch.int64)]
    if bool((listed <= 0).any()):
        raise ValueError("every streamed row needs at least one edge (A_hat rows hold their self loop)")

    n_rows_total = n
    if order is None:
        L = deg
        a = ip[:-1]
        rows_in_order = None
        stream_cols = indices.to(torch.int32)
        stream_vals = vals
    else:
        order = order.to(device=dev, dtype=torch.int64)
        if not subset and order.numel() != n:
            raise ValueError("order must list every row once (pass subset=True for a partial stream)")
        L = deg[order]
        a = torch.cumsum(L, 0) - L
        nnz = int(L.sum().item())
        n = int(order.numel())
        # source position of every stream edge
        src = torch.repeat_interleave(ip[:-1][order] - a, L) + torch.arange(nnz, device=dev, dtype=torch.int64)
        stream_cols = indices[src].to(torch.int32)
        stream_vals = None if vals is None else vals[src]
        rows_in_order = order
        del src
    b = a + L
    ca = torch.div(a, W, rounding_mode="floor")
    cb = torch.div(b - 1, W, rounding_mode="floor")
    pieces = cb - ca + 1
    n_segs = int(pieces.sum().item())
    split = pieces > 1

    if n_segs == n:  # nothing is cut
        owner = torch.arange(n, device=dev, dtype=torch.int64)
        seg_start = a
        seg_end = b - 1
        seg_partial = torch.zeros(n, dtype=torch.bool, device=dev)
    else:
        owner = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), pieces)
        first = torch.cumsum(pieces, 0) - pieces
        q = torch.arange(n_segs, device=dev, dtype=torch.int64) - first[owner]
        cq = ca[owner] + q
        seg_start = torch.maximum(a[owner], cq * W)
        seg_end = torch.minimum(b[owner], (cq + 1) * W) - 1
        seg_partial = split[owner]
        del first, q, cq

    owner_row = owner if rows_in_order is None else rows_in_order[owner]
    slot = torch.cumsum(seg_partial.to(torch.int64), 0) - 1
    n_slots = int(seg_partial.sum().item())
    seg_row = torch.where(seg_partial, slot + FLAG_I32, owner_row).to(torch.int32)
    # 64 spare entries: the kernel prefetches seg_row[s + lane] without a bounds check
    seg_row = torch.cat([seg_row, torch.zeros(64, dtype=torch.int32, device=dev)])
    n_real_segs = n_segs

    n_chunks = (nnz + W - 1) // W
    n_chunks = ((n_chunks + 31) // 32) * 32
    total = n_chunks * W
    # padding edges (after the last segment end) point at row 0 and are never emitted
    cols = torch.zeros(total, dtype=torch.int32, device=dev)
    cols[:nnz] = stream_cols
    cols[seg_end] |= FLAG_I32
    svals = None
    if stream_vals is not None:
        svals = torch.zeros(total, dtype=torch.float32, device=dev)
        svals[:nnz] = stream_vals.to(torch.float32)
    chunk_starts = torch.arange(n_chunks, device=dev, dtype=torch.int64) * W
    chunk_seg = torch.searchsorted(seg_start.contiguous(), chunk_starts, right=False).to(torch.int32)

    fix_rows_pos = torch.nonzero(split).flatten()
    fix_row = (fix_rows_pos if rows_in_order is None else rows_in_order[fix_rows_pos]).to(torch.int32)
    fp = torch.zeros(fix_rows_pos.numel() + 1, dtype=torch.int64, device=dev)
    if fix_rows_pos.numel():
        fp[1:] = torch.cumsum(pieces[fix_rows_pos], 0)
    fix_deg = L[fix_rows_pos].to(torch.float32)
    if row_deg is not None:   # partial-row stream: the epilogue needs the degree of the whole row
        row_deg = row_deg.to(device=dev, dtype=torch.float32).contiguous()
        fix_deg = row_deg[fix_row.to(torch.int64)]
    return StreamPlan(n=n_rows_total, nnz=nnz, chunk_edges=W, n_chunks=n_chunks, cols=cols, vals=svals, seg_row=seg_row,
                      chunk_seg=chunk_seg, fix_ptr=fp.to(torch.int32), fix_row=fix_row, fix_deg=fix_deg,
                      n_slots=n_slots, order=order, n_segs_real=n_real_segs, row_deg=row_deg)


def degree_order(indptr):
    """Processing order: rows by descending degree, ties by row id (stable)."""
    ip = indptr.to(torch.int64)
    deg = ip[1:] - ip[:-1]
    return torch.sort(deg, descending=True, stable=True).indices


def _plan_from_runs(n, run_row, run_len, stream_cols, stream_vals, chunk_edges, full_deg, rank, order=None):
    """Generic stream builder: ``run_row[i]`` / ``run_len[i]`` list, in stream order, runs of edges that
    belong to one row (a row may own several runs anywhere in the stream).  Runs are cut at chunk
    boundaries; a row whose edges end up in more than one segment gets one partial slot per segment
    (slots of a row are contiguous, rows in ``rank`` order) and is finished by the fix-up kernel."""
    dev = stream_cols.device
    W = chunk_edges
    nnz = int(stream_cols.numel())
    L = run_len.to(torch.int64)
    a = torch.cumsum(L, 0) - L
    b = a + L
    ca = torch.div(a, W, rounding_mode="floor")
    cb = torch.div(b - 1, W, rounding_mode="floor")
    pieces = cb - ca + 1
    n_runs = int(L.numel())
    n_segs = int(pieces.sum().item())
    owner = torch.repeat_interleave(torch.arange(n_runs, device=dev, dtype=torch.int64), pieces)
    first = torch.cumsum(pieces, 0) - pieces
    q = torch.arange(n_segs, device=dev, dtype=torch.int64) - first[owner]
    cq = ca[owner] + q
    seg_start = torch.maximum(a[owner], cq * W)
    seg_end = torch.minimum(b[owner], (cq + 1) * W) - 1
    seg_owner_row = run_row.to(torch.int64)[owner]
    del first, q, cq, owner

    per_row = torch.bincount(seg_owner_row, minlength=n)
    if bool((per_row == 0).any()):
        raise ValueError("every row needs at least one edge in the stream")
    split_row = per_row > 1
    # fix rows in rank order (heaviest first); slots of a row are contiguous
    fix_rows = torch.nonzero(split_row).flatten()
    fix_rows = fix_rows[torch.argsort(rank[fix_rows], stable=True)]
    n_fix = int(fix_rows.numel())
    fp = torch.zeros(n_fix + 1, dtype=torch.int64, device=dev)
    if n_fix:
        fp[1:] = torch.cumsum(per_row[fix_rows], 0)
    n_slots = int(fp[-1].item())
    slot_base = torch.full((n,), -1, dtype=torch.int64, device=dev)
    slot_base[fix_rows] = fp[:-1]
    # index of every segment among the segments of its row, in stream order
    by_row = torch.argsort(seg_owner_row, stable=True)
    row_first = torch.cumsum(per_row, 0) - per_row
    k_in_row = torch.empty(n_segs, dtype=torch.int64, device=dev)
    k_in_row[by_row] = torch.arange(n_segs, device=dev, dtype=torch.int64) - row_first[seg_owner_row[by_row]]
    seg_partial = split_row[seg_owner_row]
    slot = slot_base[seg_owner_row] + k_in_row
    seg_row = torch.where(seg_partial, slot + FLAG_I32, seg_owner_row).to(torch.int32)
    seg_row = torch.cat([seg_row, torch.zeros(64, dtype=torch.int32, device=dev)])

    n_chunks = (nnz + W - 1) // W
    n_chunks = ((n_chunks + 31) // 32) * 32
    total = n_chunks * W
    cols = torch.zeros(total, dtype=torch.int32, device=dev)
    cols[:nnz] = stream_cols.to(torch.int32)
    cols[seg_end] |= FLAG_I32
    svals = None
    if stream_vals is not None:
        svals = torch.zeros(total, dtype=torch.float32, device=dev)
        svals[:nnz] = stream_vals.to(torch.float32)
    chunk_starts = torch.arange(n_chunks, device=dev, dtype=torch.int64) * W
    chunk_seg = torch.searchsorted(seg_start.contiguous(), chunk_starts, right=False).to(torch.int32)
    return StreamPlan(n=n, nnz=nnz, chunk_edges=W, n_chunks=n_chunks, cols=cols, vals=svals, seg_row=seg_row,
                      chunk_seg=chunk_seg, fix_ptr=fp.to(torch.int32), fix_row=fix_rows.to(torch.int32),
                      fix_deg=full_deg[fix_rows].to(torch.float32), n_slots=n_slots, order=order,
                      n_segs_real=n_segs)


def column_ranks(indptr, indices, n_cols=None):
    """Rank of every column by how many stored entries reference it (0 = hottest; ties by column id).  For the
    symmetric A_hat that is the degree rank of the row."""
    dev = indices.device
    n = int(indptr.numel()) - 1
    m_cols = n if n_cols is None else int(n_cols)
    if n_cols is None:
        ip = indptr.to(torch.int64)
        refs = ip[1:] - ip[:-1]
    else:
        refs = torch.bincount(indices.to(torch.int64), minlength=m_cols)
    corder = torch.sort(refs, descending=True, stable=True).indices
    crank = torch.empty(m_cols, dtype=torch.int64, device=dev)
    crank[corder] = torch.arange(m_cols, device=dev, dtype=torch.int64)
    return crank


def rank_sorted_csr(indptr, indices, vals=None, n_cols=None, max_batch_edges=1 << 27):
    """The same CSR with the entries of every row sorted by the rank of their column (hottest first) instead of
    the column id.  Which (row, column, value) triples exist does not change; only the order in which a row's
    gathers are issued and added does.  Sorted in row blocks of at most ``max_batch_edges`` entries so that the
    temporaries stay small next to a multi-GB shard.  Returns (indices, vals, crank)."""
    dev = indices.device
    n = int(indptr.numel()) - 1
    ip = indptr.to(torch.int64)
    crank = column_ranks(indptr, indices, n_cols)
    m_cols = int(crank.numel())
    out_idx = torch.empty_like(indices)
    out_val = None if vals is None else torch.empty_like(vals)
    nnz = int(ip[-1].item())
    r0 = 0
    while r0 < n:
        # rows [r0, r1): as many whole rows as fit the batch (at least one)
        lim = int(ip[r0].item()) + int(max_batch_edges)
        r1 = int(torch.searchsorted(ip, torch.tensor([lim], device=dev, dtype=torch.int64), right=True).item()) - 1
        r1 = min(max(r1, r0 + 1), n)
        e0, e1 = int(ip[r0].item()), int(ip[r1].item())
        if e1 > e0:
            row_of = torch.repeat_interleave(torch.arange(r1 - r0, device=dev, dtype=torch.int64), ip[r0 + 1: r1 + 1] - ip[r0: r1])
            perm = torch.sort(row_of * m_cols + crank[indices[e0:e1].to(torch.int64)], stable=True).indices
            out_idx[e0:e1] = indices[e0:e1][perm]
            if vals is not None:
                out_val[e0:e1] = vals[e0:e1][perm]
            del row_of, perm
        r0 = r1
    assert nnz == int(indices.numel())
    return out_idx, out_val, crank


def permute_chunks(plan: StreamPlan, perm):
    """Chunks are independent work items (``chunk_seg[c]`` points at a chunk's segments wherever the chunk sits):
    process them in the order ``perm`` (new position -> old chunk).  Only cols / vals / chunk_seg move."""
    if plan.lane_group:
        raise ValueError("permute the chunks before the lane transposition")
    nc, W = plan.n_chunks, plan.chunk_edges
    perm = perm.to(device=plan.cols.device, dtype=torch.int64)
    if perm.numel() != nc:
        raise ValueError("perm must list every chunk once")
    cols = plan.cols.view(nc, W)[perm].contiguous().view(-1)
    vals = None if plan.vals is None else plan.vals.view(nc, W)[perm].contiguous().view(-1)
    return StreamPlan(plan.n, plan.nnz, W, nc, cols, vals, plan.seg_row, plan.chunk_seg[perm].contiguous(), plan.fix_ptr,
                      plan.fix_row, plan.fix_deg, plan.n_slots, plan.order, plan.n_segs_real, plan.row_deg,
                      plan.wide_cta, plan.lane_group, plan.carve)


def window_order_chunks(plan: StreamPlan, crank, key="first"):
    """Process the chunks that hold ONE whole segment (the interior chunks of the hub rows, each already a partial
    sum with its own slot) sorted by the rank of their ``key`` column (first / mid / last of the chunk) -- with
    rank-sorted rows (``rank_sorted_csr``) all SMs then sweep the column space hot end first and together, so the
    rows one warp gathers are L1 / L2 hits for the others.  No new partial sums: only whole chunks move.  The
    remaining chunks follow in their old order (merging them into the sweep at their rows' rank was modelled,
    tools/window_model.py history, and loses: it stretches the hub sweep)."""
    nc, W = plan.n_chunks, plan.chunk_edges
    dev = plan.cols.device
    C = plan.cols.view(nc, W)
    single = (C[:, :-1] >= 0).all(dim=1) & (C[:, -1] < 0)
    pos = {"first": 0, "mid": W // 2, "last": W - 1}[key]
    kcol = (C[:, pos] & 0x7FFFFFFF).to(torch.int64)
    crank = crank.to(dev)
    big = int(crank.numel())
    k = torch.where(single, crank[kcol], big + torch.arange(nc, device=dev, dtype=torch.int64))
    return permute_chunks(plan, torch.sort(k, stable=True).indices)


def interleave_chunks(plan: StreamPlan, n_lead_chunks, unit_chunks=64):
    """Permute the chunks of ``plan`` so that units of ``unit_chunks`` consecutive chunks of its leading part
    (the first ``n_lead_chunks`` chunks: the carved pieces, served from L1) alternate evenly with units of
    the rest (the residual rows, whose cold gathers go to DRAM).  Chunks are independent work items --
    ``chunk_seg[c]`` points at a chunk's segments wherever the chunk sits -- so only cols / vals / chunk_seg
    move.  Measured reason (profiles/r01_variants.md): streamed back to back, the carved part leaves HBM idle
    and the residual part then runs at the random-access limit of the DRAM; interleaved, every SM wave
    mixes both."""
    nc = plan.n_chunks
    lead = max(0, min(int(n_lead_chunks), nc))
    if lead == 0 or lead == nc:
        return plan
    dev = plan.cols.device
    u = int(unit_chunks)
    n_lead_units = (lead + u - 1) // u
    n_rest_units = (nc - lead + u - 1) // u
    # merge the two unit sequences by their fractional position
    key = torch.cat([(torch.arange(n_lead_units, dtype=torch.float64) + 0.5) / n_lead_units,
                     (torch.arange(n_rest_units, dtype=torch.float64) + 0.5) / n_rest_units])
    start = torch.cat([torch.arange(n_lead_units, dtype=torch.int64) * u, lead + torch.arange(n_rest_units, dtype=torch.int64) * u])
    stop = torch.cat([torch.clamp(start[:n_lead_units] + u, max=lead), torch.clamp(start[n_lead_units:] + u, max=nc)])
    o = torch.argsort(key, stable=True)
    start, stop = start[o], stop[o]
    lens = stop - start
    off = torch.cumsum(lens, 0) - lens
    perm = (torch.repeat_interleave(start - off, lens) + torch.arange(nc, dtype=torch.int64)).to(dev)
    W = plan.chunk_edges
    cols = plan.cols.view(nc, W)[perm].contiguous().view(-1)
    vals = None if plan.vals is None else plan.vals.view(nc, W)[perm].contiguous().view(-1)
    out = StreamPlan(plan.n, plan.nnz, W, nc, cols, vals, plan.seg_row, plan.chunk_seg[perm].contiguous(), plan.fix_ptr,
                     plan.fix_row, plan.fix_deg, plan.n_slots, plan.order, plan.n_segs_real, plan.row_deg,
                     plan.wide_cta, plan.lane_group, plan.carve)
    return out


def build_carved_plan(indptr, indices, vals=None, chunk_edges=256, block_cols=512, n_blocks=64, min_piece=4,
                      wide_cta=True, interleave=False, unit_chunks=64, n_cols=None, levels=None):
    """Edge stream with the hot column blocks carved out ("carved stream").

    Columns are ranked by degree; block b holds the columns of rank [b * block_cols, (b+1) * block_cols)
    for b < n_blocks -- ``block_cols`` rows of the feature matrix are meant to fit one SM's L1.  The
    edges of a row that fall into one block form a *piece*; pieces of at least ``min_piece`` edges are
    taken out of their row and streamed block by block (rows in degree order inside a block), so
    that an SM walking consecutive chunks keeps gathering the same ``block_cols`` rows.  What is left
    of every row follows in degree order, exactly as in ``build_stream_plan(order=degree_order)``.
    Every carved row becomes a split row (partial sums + fix-up), results are order-deterministic.
    Column ids are not relabelled.  ``interleave=True`` alternates units of ``unit_chunks`` carved chunks
    with units of residual chunks (``interleave_chunks``).  ``n_cols``: width of the column space when it is
    not the row count (a shard of the partitioned form: local rows x [local | halo] columns); columns are
    ranked by how many stored entries reference them -- for the symmetric A_hat that IS the row degree.
    With blocks sized for the L2 instead of the L1 (e.g. 16 blocks of n/16 columns) the same stream keeps the
    cold gathers of the hub rows inside an L2-resident window (tools/l1sim.c with one cache models it).
    ``levels`` = [(block_cols, n_blocks, min_piece), ...] stacks several block sizes along the rank axis
    (e.g. 64 L1-sized blocks over the hottest columns, then L2-sized blocks over the rest) and replaces
    the three scalar parameters.
    """
    if chunk_edges % 128 != 0 or chunk_edges <= 0:
        raise ValueError("chunk_edges must be a positive multiple of 128")
    if levels is None:
        levels = [(block_cols, n_blocks, min_piece)]
    for bc_, nb_, t_ in levels:
        if bc_ <= 0 or nb_ < 0 or t_ < 1:
            raise ValueError("block_cols > 0, n_blocks >= 0, min_piece >= 1")
    dev = indices.device
    n = int(indptr.numel()) - 1
    ip = indptr.to(torch.int64)
    nnz = int(ip[-1].item())
    if nnz != int(indices.numel()):
        raise ValueError("indptr[-1] != len(indices)")
    deg = ip[1:] - ip[:-1]
    if bool((deg <= 0).any()):
        raise ValueError("every streamed row needs at least one edge (A_hat rows hold their self loop)")
    order = degree_order(indptr)
    rank = torch.empty(n, dtype=torch.int64, device=dev)
    rank[order] = torch.arange(n, device=dev, dtype=torch.int64)
    m_cols = n if n_cols is None else int(n_cols)
    if m_cols == n and n_cols is None:
        crank = rank                     # symmetric pattern: references per column == row degree
    else:
        if nnz and int(indices.max().item()) >= m_cols:
            raise ValueError("a column index lies outside n_cols")
        refs = torch.bincount(indices.to(torch.int64), minlength=m_cols)
        corder = torch.sort(refs, descending=True, stable=True).indices
        crank = torch.empty(m_cols, dtype=torch.int64, device=dev)
        crank[corder] = torch.arange(m_cols, device=dev, dtype=torch.int64)
        del refs, corder
    ends, tmins = [], []                 # block b = ranks [ends[b-1], ends[b]), carved from pieces of >= tmins[b] edges
    for bc_, nb_, t_ in levels:
        for _ in range(int(nb_)):
            lo = ends[-1] if ends else 0
            if lo >= m_cols:
                break
            ends.append(min(lo + int(bc_), m_cols))
            tmins.append(int(t_))
    NB = len(ends)
    row_of = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), deg)
    blk = torch.bucketize(crank[indices.to(torch.int64)], torch.tensor(ends, dtype=torch.int64, device=dev), right=True)
    tmin = torch.tensor(tmins + [1 << 62], dtype=torch.int64, device=dev)
    # edges per (row, block): CSR order is row-major, so a stable sort by block inside the row groups them
    key = row_of * (NB + 1) + blk
    skey, perm = torch.sort(key, stable=True)
    _, inv, cnt = torch.unique_consecutive(skey, return_inverse=True, return_counts=True)
    keep = (cnt[inv] >= tmin[skey % (NB + 1)]) & ((skey % (NB + 1)) < NB)
    fblk = torch.full((nnz,), NB, dtype=torch.int64, device=dev)
    fblk[perm] = torch.where(keep, skey % (NB + 1), torch.full_like(skey, NB))
    carved_edges = int(keep.sum().item())
    del key, skey, inv, cnt, keep, blk
    # stream order: (block, rank of the row), columns ascending inside a run (stable sort of CSR order)
    key2 = fblk * n + rank[row_of]
    skey2, perm2 = torch.sort(key2, stable=True)
    run_key, run_len = torch.unique_consecutive(skey2, return_counts=True)
    run_row = order[run_key % n]
    stream_cols = indices[perm2]
    stream_vals = None if vals is None else vals[perm2]
    n_carved_runs = int((run_key < NB * n).sum().item())
    del key2, skey2, perm2, fblk, row_of
    plan = _plan_from_runs(n, run_row, run_len, stream_cols, stream_vals, chunk_edges, deg, rank, order)
    plan.wide_cta = bool(wide_cta)
    plan.carve = {"block_cols": levels[0][0], "n_blocks": NB, "min_piece": levels[0][2],
                  "levels": [list(map(int, lv)) for lv in levels], "carved_edges": carved_edges,
                  "carved_pieces": n_carved_runs, "n_slots": plan.n_slots, "n_fix": plan.n_fix,
                  "interleave": bool(interleave)}
    if interleave:
        plan = interleave_chunks(plan, carved_edges // chunk_edges, unit_chunks)
    return plan


def lane_group_for(F):
    """Lane group the propagation kernel uses for feature width F (csrc/appnp_spmm.cu dispatch_step)."""
    g = 1
    w = F // 4 if F % 4 == 0 else F
    while g < w:
        g <<= 1
    return min(g, 32)


def lane_transpose(plan: StreamPlan, G):
    """Copy of ``plan`` whose cols / vals are stored lane-transposed for lane groups of G lanes
    (include/ppnp_b200.h PPNP_PLAN_LANE_GROUP): each lane's 4 index words of 4/SR consecutive slabs
    are contiguous, so the kernel stages them with one 16-byte cp.async per lane."""
    if G not in (4, 8, 16, 32):
        raise ValueError("lane-transposed streams exist for lane groups of 4, 8, 16 or 32")
    if plan.lane_group:
        raise ValueError("plan is already lane-transposed")
    W = plan.chunk_edges
    SR = max(1, 16 // G)
    SE = SR * G
    CPS = 4 // SR
    if (W // SE) % CPS != 0:
        raise ValueError("chunk_edges does not hold whole staging quads for this lane group")
    dev = plan.cols.device
    p = torch.arange(W, device=dev, dtype=torch.int64)
    j, r, l = p // SE, (p % SE) // G, p % G
    stored = (j // CPS) * (CPS * SE) + l * 4 + (j % CPS) * SR + r
    src = torch.empty(W, dtype=torch.int64, device=dev)
    src[stored] = p                                   # stored position -> logical position
    cols = plan.cols.view(plan.n_chunks, W)[:, src].contiguous().view(-1)
    vals = None if plan.vals is None else plan.vals.view(plan.n_chunks, W)[:, src].contiguous().view(-1)
    return StreamPlan(plan.n, plan.nnz, W, plan.n_chunks, cols, vals, plan.seg_row, plan.chunk_seg, plan.fix_ptr,
                      plan.fix_row, plan.fix_deg, plan.n_slots, plan.order, plan.n_segs_real, plan.row_deg,
                      plan.wide_cta, G, plan.carve)
