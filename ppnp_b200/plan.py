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

import torch

from . import _lib

FLAG_I32 = -(1 << 31)          # PPNP_FLAG as a signed int32


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
            s.chunk_edges, s.reserved = self.chunk_edges, 0
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
                          mv(self.fix_deg), self.n_slots, mv(self.order), self.n_segs_real, mv(self.row_deg))


def build_stream_plan(indptr, indices, vals=None, chunk_edges=256, order=None, subset=False, row_deg=None):
    """Cut the CSR (indptr[n+1], indices[nnz], optional vals[nnz]) into the edge stream.

    ``order`` (int64, a permutation of the rows) is the processing order; column ids are never
    relabelled, so inputs and outputs of the propagation keep the caller's row order.  With
    ``subset=True`` ``order`` may list only some rows: the stream then produces exactly those rows
    (used to split a shard into interior and boundary rows, ppnp_b200/dist.py).
    """
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
