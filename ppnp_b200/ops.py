"""Host-side operators over libppnp_b200.so.  Tensors in, tensors out; every op runs on the
current CUDA stream and raises when the CUDA library is missing (no fallback path).

Reference lines each operator replaces are cited per function (paths into the reference tree).
"""
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .plan import (StreamPlan, build_carved_plan, build_stream_plan, degree_order, lane_group_for, lane_transpose,
                   rank_sorted_csr, window_order_chunks)
from .tiled import TiledPlan, build_tiled_plan

MODE = {"sym": _lib.MODE_SYM, "rw": _lib.MODE_RW}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ppnp_b200 ops need CUDA tensors (there is no CPU path)")


@dataclass
class NormalizedCSR:
    """A_hat of helpers.py:58-66 on the device: structure of adj + I, D, values, D^-1/2."""
    n: int
    nnz: int
    indptr: torch.Tensor     # int32 [n+1]
    indices: torch.Tensor    # int32 [nnz]
    deg: torch.Tensor        # fp64 [n]   D = rowsum(adj + I)
    val64: Optional[torch.Tensor]  # fp64 [nnz]
    val32: Optional[torch.Tensor]  # fp32 [nnz]
    dinv: torch.Tensor       # fp32 [n]
    mode: str
    unit_weights: bool = True  # adj was all ones with an empty diagonal: D_i == number of stored entries of row i of A_hat


def csr_normalize(indptr, indices, data=None, mode="sym", want_val64=False, want_val32=True):
    """helpers.py:58-66 ``calc_A_hat(adj, mode)`` on the GPU.

    indptr/indices: canonical CSR of ``adj`` (int32, sorted, CUDA tensors); data: fp32 weights or
    None for the all-ones adjacency that ``SparseGraph.standardize`` produces.
    """
    lib = _lib.load()
    _require_cuda(indptr, indices, data)
    indptr = indptr.to(torch.int32).contiguous()
    indices = indices.to(torch.int32).contiguous()
    if data is not None:
        data = data.to(torch.float32).contiguous()
    n = indptr.numel() - 1
    nnz = indices.numel()
    dev = indices.device
    cap = nnz + n
    out_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    out_indices = torch.empty(cap, dtype=torch.int32, device=dev)
    out_deg = torch.empty(n, dtype=torch.float64, device=dev)
    out_val64 = torch.empty(cap, dtype=torch.float64, device=dev) if want_val64 else None
    out_val32 = torch.empty(cap, dtype=torch.float32, device=dev) if want_val32 else None
    out_dinv = torch.empty(n, dtype=torch.float32, device=dev)
    ws_bytes = lib.ppnp_csr_normalize_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.ppnp_csr_normalize(_lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data), n, nnz, MODE[mode],
                                    _lib.ptr(out_indptr), _lib.ptr(out_indices), _lib.ptr(out_deg),
                                    _lib.ptr(out_val64), _lib.ptr(out_val32), _lib.ptr(out_dinv),
                                    _lib.ptr(ws), ws_bytes, _lib.current_stream())
    _lib.check(rc, "ppnp_csr_normalize")
    nnz_hat = int(out_indptr[-1].item())
    # the value-free propagation (csrc/appnp_spmm.cu) takes a row's degree from its edge count: that equals
    # D = rowsum(adj + I) only for unit weights and a diagonal that was empty (every row gained exactly one entry)
    unit = data is None and nnz_hat == nnz + n
    return NormalizedCSR(n=n, nnz=nnz_hat, indptr=out_indptr, indices=out_indices[:nnz_hat], deg=out_deg,
                         val64=None if out_val64 is None else out_val64[:nnz_hat],
                         val32=None if out_val32 is None else out_val32[:nnz_hat], dinv=out_dinv, mode=mode,
                         unit_weights=bool(unit))


class PropagationGraph:
    """Normalised adjacency + edge-stream plan, ready for ``appnp_propagate``."""

    def __init__(self, ahat: NormalizedCSR, chunk_edges=256, order="natural", keep_vals=True, idx16=False, carve=None,
                 tiled=None, rows_below=None, window=None):
        """order: "natural" | "degree" | a permutation tensor (row-major streams, plan.build_stream_plan),
        "carve" (hot column blocks first, plan.build_carved_plan; ``carve`` = its keyword arguments) or
        "window" (degree order with every row's columns sorted by rank and the whole-segment chunks of the hub
        rows processed in column-window order, plan.window_order_chunks; ``window`` = {"key": "first" | "mid" |
        "last", "wide_cta": bool}).
        idx16: stage the index stream with 16-byte copies from a lane-transposed copy of the stream
        (feature widths 16 and 64; other widths keep the linear stream).
        tiled: keyword arguments of ``tiled.build_tiled_plan`` plus ``slice_width`` (floats of the feature
        dimension per CTA: 16, 32 or 64): rows of high degree run through the shared-memory-resident kernel
        (csrc/appnp_tiled.cu), the rest through the row-major stream.  Built per feature width on first use.
        rows_below: rows with fewer stored entries than this go through the one-lane-group-per-row kernel
        (csrc/appnp_rows.cu) straight off the CSR; the edge stream (or the tiled plan) then holds only the rows of
        higher degree.  Needs unit weights unless the stored values are kept."""
        self.ahat = ahat
        self.unit_weights = bool(getattr(ahat, "unit_weights", True))
        if not self.unit_weights and not keep_vals:
            raise ValueError("a weighted adjacency (or one with stored diagonal entries) needs the stored values: keep_vals=True")
        if not self.unit_weights:
            tiled = None            # the tiled plan's epilogue takes degrees from edge counts as well
        self.tiled_kw = None if tiled is None else dict(tiled)
        self._tiled = {}
        self._keep_vals = bool(keep_vals)
        self._chunk_edges = int(chunk_edges)
        self.mode = ahat.mode
        vals = ahat.val32 if keep_vals else None
        self.rows_part = None       # (rows int32 desc-degree list, RowsPlanStruct) of the low-degree rows
        self._hub_order = None
        if rows_below is not None:
            if isinstance(order, str) and order == "carve":
                raise ValueError("rows_below goes with the row-major orders")
            full = degree_order(ahat.indptr)
            ipl = ahat.indptr.to(torch.int64)
            n_hub = int(((ipl[1:] - ipl[:-1]) >= int(rows_below)).sum().item())
            self._hub_order = full[:n_hub]
            low = full[n_hub:].to(torch.int32).contiguous()
            if low.numel():
                rs = _lib.RowsPlanStruct()
                rs.n, rs.n_rows = ahat.n, int(low.numel())
                rs.indptr, rs.indices = ahat.indptr.data_ptr(), ahat.indices.data_ptr()
                rs.vals = vals.data_ptr() if vals is not None else None
                rs.rows = low.data_ptr()
                self.rows_part = (low, rs)
        if self._hub_order is not None:
            self.plan = (build_stream_plan(ahat.indptr, ahat.indices, vals, chunk_edges, self._hub_order, subset=True)
                         if self._hub_order.numel() else None)
        elif isinstance(order, str) and order == "carve":
            self.plan: StreamPlan = build_carved_plan(ahat.indptr, ahat.indices, vals, chunk_edges, **(carve or {}))
        elif isinstance(order, str) and order == "window":
            wkw = dict(window or {})
            sidx, svals, crank = rank_sorted_csr(ahat.indptr, ahat.indices, vals)
            base = build_stream_plan(ahat.indptr, sidx, svals, chunk_edges, degree_order(ahat.indptr))
            self.plan = window_order_chunks(base, crank, key=wkw.get("key", "mid"))
            self.plan.wide_cta = bool(wkw.get("wide_cta", False))
            del sidx, svals, crank, base
        else:
            if carve is not None:
                raise ValueError("carve parameters need order='carve'")
            if isinstance(order, str) and order == "natural":
                ord_t = None
            elif isinstance(order, str) and order == "degree":
                ord_t = degree_order(ahat.indptr)
            elif torch.is_tensor(order):
                ord_t = order
            else:
                raise ValueError(f"unknown order {order!r}")
            self.plan = build_stream_plan(ahat.indptr, ahat.indices, vals, chunk_edges, ord_t)
        self.n = ahat.n
        self.nnz = ahat.nnz
        self.idx16 = bool(idx16)
        self._plans16 = {}
        self._partial = {}

    def plan_for(self, F):
        """The stream a propagation over F features reads (lane-transposed copy when idx16 applies)."""
        if self.plan is None or not self.idx16 or F not in (16, 64):
            return self.plan
        G = lane_group_for(F)
        if G not in self._plans16:
            self._plans16[G] = lane_transpose(self.plan, G)
        return self._plans16[G]

    def tiled_for(self, F):
        """(TiledPlan, row-major plan of the remaining rows or None, slice width) for feature width F, or None when
        the graph was not built with ``tiled=`` or F is not a multiple of the slice width."""
        if self.tiled_kw is None:
            return None
        kw = dict(self.tiled_kw)
        W = int(kw.pop("slice_width", 64))
        if F % W != 0 or F % 4 != 0:
            return None
        if F not in self._tiled:
            lib = _lib.load()
            if "n_ctas" not in kw:
                import ctypes
                sm = ctypes.c_int32(0)
                _lib.check(lib.ppnp_device_info(ctypes.byref(sm), None, None, None), "ppnp_device_info")
                kw["n_ctas"] = max(1, sm.value // (F // W))
            kw.setdefault("slot_rows", (100 * 1024 - 1024 - 128) // (W * 4) - 1)
            kw.setdefault("rest_chunk_edges", self._chunk_edges)
            rows_rest = kw.pop("rest", "stream") == "rows"
            vals = self.ahat.val32 if self._keep_vals else None
            tp = build_tiled_plan(self.ahat.indptr, self.ahat.indices, vals, build_rest=not rows_rest, **kw)
            rest = tp.rest
            if rest is not None and self.idx16 and F in (16, 64):
                rest = lane_transpose(rest, lane_group_for(F))
            rows = None
            if rows_rest:       # every row the tiled plan does not own goes to the one-group-per-row kernel, by descending degree
                full = degree_order(self.ahat.indptr)
                low = full[tp.stats["hub_rows"]:].to(torch.int32).contiguous()
                if low.numel():
                    rs = _lib.RowsPlanStruct()
                    rs.n, rs.n_rows = self.ahat.n, int(low.numel())
                    rs.indptr, rs.indices = self.ahat.indptr.data_ptr(), self.ahat.indices.data_ptr()
                    rs.vals = vals.data_ptr() if vals is not None else None
                    rs.rows = low.data_ptr()
                    rows = (low, rs)
            self._tiled[F] = (tp, rest, W, rows)
        return self._tiled[F]

    def parts_for(self, F):
        """(tiled plan | None, stream plan | None, rows part | None, slice width, partial buffer of the stream part):
        the kernels a step over F features runs, together producing every row once."""
        tl = self.tiled_for(F)
        if tl is not None:
            tp, rest, W, rows = tl
            return tp, rest, rows, W, self.rest_partial_buffer(rest, F)
        return None, self.plan_for(F), self.rows_part, 64, self.partial_buffer(F)

    @classmethod
    def from_adjacency(cls, indptr, indices, data=None, mode="sym", **kw):
        return cls(csr_normalize(indptr, indices, data, mode), **kw)

    def partial_buffer(self, ld):
        if self.plan is None or self.plan.n_slots == 0:
            return None
        buf = self._partial.get(ld)
        if buf is None:
            buf = torch.empty(self.plan.n_slots * ld, dtype=torch.float32, device=self.plan.device)
            self._partial[ld] = buf
        return buf

    def rest_partial_buffer(self, rest, ld):
        if rest is None or rest.n_slots == 0:
            return None
        key = ("rest", ld)
        buf = self._partial.get(key)
        if buf is None:
            buf = torch.empty(rest.n_slots * ld, dtype=torch.float32, device=rest.device)
            self._partial[key] = buf
        return buf


def spmm_step(graph: PropagationGraph, Zin, T, alpha, epi=_lib.EPI_PLAIN, use_vals=True, out=None):
    """One propagation step  out = a * A Zin + b * T  (epilogue ``epi``, include/ppnp_b200.h)."""
    lib = _lib.load()
    _require_cuda(Zin, T)
    Zin = Zin.contiguous()
    T = T.contiguous()
    n, F = Zin.shape
    if not graph.unit_weights and (not use_vals or (int(epi) & 15) != _lib.EPI_PLAIN):
        raise ValueError("this graph has weights or stored diagonal entries: only the stored-value step (use_vals=True, "
                         "EPI_PLAIN) is defined for it -- the value-free epilogues take degrees from edge counts")
    if out is None:
        out = torch.empty_like(Zin)
    tiled, stream_plan, rows, W, partial = graph.parts_for(F)
    with torch.cuda.device(Zin.device):
        if tiled is not None:
            rc = lib.ppnp_spmm_step_tiled(tiled.struct(), _lib.ptr(Zin), _lib.ptr(T), _lib.ptr(out), F, F, W, float(alpha), int(epi),
                                          int(bool(use_vals)), _lib.current_stream())
            _lib.check(rc, "ppnp_spmm_step_tiled")
        if stream_plan is not None:
            rc = lib.ppnp_spmm_step(stream_plan.struct(), _lib.ptr(Zin), _lib.ptr(T), _lib.ptr(out), _lib.ptr(partial),
                                    F, F, float(alpha), int(epi), int(bool(use_vals)), _lib.current_stream())
            _lib.check(rc, "ppnp_spmm_step")
        if rows is not None:
            rs = rows[1]
            rc = lib.ppnp_spmm_step_rows(rs.indptr, rs.indices, rs.vals, rs.rows, rs.n_rows, rs.n, _lib.ptr(Zin), _lib.ptr(T), _lib.ptr(out),
                                         F, F, float(alpha), int(epi), int(bool(use_vals)), None, None, None, None, 0, _lib.current_stream())
            _lib.check(rc, "ppnp_spmm_step_rows")
    return out


def appnp_propagate(graph: PropagationGraph, H, K=10, alpha=0.1, use_vals=False, out=None, scratch=None, scaled_input=False,
                    per_step=False):
    """K steps of Z <- (1-alpha) A_hat Z + alpha H from Z_0 = H (north_star; forward and backward).

    ``use_vals=False`` runs the value-free Y-space iteration (stored values only in step 1),
    ``use_vals=True`` multiplies by the stored A_hat values in every step.
    ``scaled_input=True``: ``H`` holds D^-1/2 H already (``linear_rowscale`` writes it that way): every step is
    value-free, the stored values are never read (PPNP_MODE_SYM_Y0); the result is the same Z.
    ``per_step=True``: K launches (+ fix-ups) even on a graph small enough for the one-launch cooperative kernel
    (PPNP_MODE_PER_STEP).
    """
    lib = _lib.load()
    _require_cuda(H)
    if H.dtype != torch.float32 or H.dim() != 2:
        raise ValueError("H must be a 2-D float32 tensor")
    H = H.contiguous()
    n, F = H.shape
    if n != graph.n:
        raise ValueError(f"H has {n} rows, graph has {graph.n}")
    if K == 0:
        return H.clone()
    if scaled_input and (graph.mode != "sym" or not graph.unit_weights or use_vals):
        raise ValueError("scaled_input is the value-free 'sym' iteration on a unit-weight graph")
    if not graph.unit_weights:
        use_vals = True     # edge counts are not degrees here: the value-free form would be silently wrong
    mode_code = (_lib.MODE_SYM_Y0 if scaled_input else MODE[graph.mode]) | (_lib.MODE_PER_STEP if per_step else 0)
    Z = out if out is not None else torch.empty_like(H)
    scratch = scratch if scratch is not None else torch.empty_like(H)
    tiled, stream_plan, rows, W, partial = graph.parts_for(F)
    if tiled is not None or rows is not None:
        with torch.cuda.device(H.device):
            rc = lib.ppnp_appnp_propagate_parts(None if tiled is None else tiled.struct(), None if stream_plan is None else stream_plan.struct(),
                                                None if rows is None else rows[1], _lib.ptr(H), _lib.ptr(Z), _lib.ptr(scratch),
                                                _lib.ptr(partial), F, F, W, int(K), float(alpha), mode_code, int(bool(use_vals)),
                                                _lib.current_stream())
        _lib.check(rc, "ppnp_appnp_propagate_parts")
        return Z
    partial = graph.partial_buffer(F)
    with torch.cuda.device(H.device):
        rc = lib.ppnp_appnp_propagate(graph.plan_for(F).struct(), _lib.ptr(H), _lib.ptr(Z), _lib.ptr(scratch),
                                      _lib.ptr(partial), F, F, int(K), float(alpha), mode_code,
                                      int(bool(use_vals)), _lib.current_stream())
    _lib.check(rc, "ppnp_appnp_propagate")
    return Z


def appnp_propagate_persistent(graph: PropagationGraph, H, K=10, alpha=0.1):
    """The K steps in one cooperative launch (csrc/appnp_spmm.cu appnp_persistent_kernel)."""
    lib = _lib.load()
    _require_cuda(H)
    H = H.contiguous()
    n, F = H.shape
    Z, scratch = torch.empty_like(H), torch.empty_like(H)
    partial = graph.partial_buffer(F)
    with torch.cuda.device(H.device):
        rc = lib.ppnp_appnp_propagate_persistent(graph.plan.struct(), _lib.ptr(H), _lib.ptr(Z), _lib.ptr(scratch),
                                                 _lib.ptr(partial), F, F, int(K), float(alpha), _lib.current_stream())
    _lib.check(rc, "ppnp_appnp_propagate_persistent")
    return Z


class _APPNPFunction(torch.autograd.Function):
    """dH = P_K(A_hat) dZ: the same K-step kernel on the upstream gradient (A_hat symmetric,
    SURVEY.md section 3.3) -- no activations are saved."""

    @staticmethod
    def forward(ctx, H, graph, K, alpha, use_vals):
        ctx.graph, ctx.K, ctx.alpha, ctx.use_vals = graph, K, alpha, use_vals
        return appnp_propagate(graph, H, K, alpha, use_vals)

    @staticmethod
    def backward(ctx, gZ):
        if ctx.graph.mode != "sym":
            raise RuntimeError("backward is implemented for the symmetric normalisation only")
        return appnp_propagate(ctx.graph, gZ.contiguous(), ctx.K, ctx.alpha, ctx.use_vals), None, None, None, None


def appnp(H, graph, K=10, alpha=0.1, use_vals=False):
    """Differentiable APPNP propagation."""
    return _APPNPFunction.apply(H, graph, K, alpha, use_vals)


# ------------------------------------------------------------------------------ fused encoder tail
def linear_rowscale(A, W, bias=None, scale=None):
    """``scale[:, None] * (A @ W.T + bias)`` (csrc/encoder_tail.cu): model.py:51's last linear with the row scaling of
    the propagation fused in.  A: [n, hidden] fp32, W: [C, hidden] (nn.Linear layout)."""
    lib = _lib.load()
    _require_cuda(A, W, bias, scale)
    A, W = A.contiguous(), W.contiguous()
    n, hidden = A.shape
    C = W.shape[0]
    if W.shape[1] != hidden:
        raise ValueError("W must be [C, hidden]")
    out = torch.empty((n, C), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        rc = lib.ppnp_linear_rowscale(_lib.ptr(A), n, hidden, _lib.ptr(W), _lib.ptr(None if bias is None else bias.contiguous()),
                                      _lib.ptr(None if scale is None else scale.contiguous()), _lib.ptr(out), C, C, _lib.current_stream())
    _lib.check(rc, "ppnp_linear_rowscale")
    return out


def linear_rowscale_backward(A, dOut, W, scale=None, need_dA=True, need_dbias=False):
    """Adjoint of ``linear_rowscale``: (dA, dW, dbias)."""
    lib = _lib.load()
    _require_cuda(A, dOut, W, scale)
    A, W, dOut = A.contiguous(), W.contiguous(), dOut.contiguous()
    n, hidden = A.shape
    C = W.shape[0]
    dA = torch.empty_like(A) if need_dA else None
    dW = torch.empty_like(W)
    dbias = torch.empty(C, dtype=torch.float32, device=A.device) if need_dbias else None
    ws_bytes = lib.ppnp_linear_rowscale_backward_workspace_bytes(n, hidden, C)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=A.device)
    with torch.cuda.device(A.device):
        rc = lib.ppnp_linear_rowscale_backward(_lib.ptr(A), _lib.ptr(dOut), C, n, hidden, C, _lib.ptr(W),
                                               _lib.ptr(None if scale is None else scale.contiguous()), _lib.ptr(dA), _lib.ptr(dW),
                                               _lib.ptr(dbias), _lib.ptr(ws), ws_bytes, _lib.current_stream())
    _lib.check(rc, "ppnp_linear_rowscale_backward")
    return dA, dW, dbias


class _FusedTailAPPNP(torch.autograd.Function):
    """Z = P_K(A_hat) (A1 @ W^T + b) without H: the last linear writes Y0 = D^-1/2 H, the K steps run value-free
    (PPNP_MODE_SYM_Y0).  Backward: P_K is symmetric, so dH = P_K(dZ) = D^1/2 P_Y(D^-1/2 dZ) through the same two
    kernels -- the scaling of dZ rides in the adjoint of the linear layer."""

    @staticmethod
    def forward(ctx, A1, W, bias, graph, K, alpha):
        dinv = graph.ahat.dinv
        Y0 = linear_rowscale(A1, W, bias, dinv)
        Z = appnp_propagate(graph, Y0, K, alpha, scaled_input=True)
        ctx.save_for_backward(A1, W)
        ctx.graph, ctx.K, ctx.alpha, ctx.has_bias = graph, K, alpha, bias is not None
        return Z

    @staticmethod
    def backward(ctx, gZ):
        A1, W = ctx.saved_tensors
        g = ctx.graph
        dinv = g.ahat.dinv
        # dH = P(gZ): scale gZ by D^-1/2 (elementwise, n x C), propagate value-free
        G0 = gZ.contiguous() * dinv[:, None]
        dH = appnp_propagate(g, G0, ctx.K, ctx.alpha, scaled_input=True)
        dA, dW, db = linear_rowscale_backward(A1, dH, W, None, need_dA=ctx.needs_input_grad[0], need_dbias=ctx.has_bias)
        return dA, dW, db, None, None, None


def appnp_fused_tail(A1, W, bias, graph, K=10, alpha=0.1):
    """Differentiable ``appnp(A1 @ W.T + bias, graph)`` with the encoder's last linear fused into the propagation's
    input scaling (no H, no stored A_hat values)."""
    if graph.mode != "sym" or not graph.unit_weights:
        raise ValueError("the fused tail is the value-free 'sym' iteration on a unit-weight graph")
    return _FusedTailAPPNP.apply(A1, W, bias, graph, K, alpha)


# ------------------------------------------------------------------------------ exact PPNP
def ppr_steps_for_tol(alpha, tol):
    """Steps for the power iteration to reach spectral-norm error ``tol``: (1-alpha)^K <= tol."""
    import math
    return max(1, int(math.ceil(math.log(tol) / math.log(1.0 - alpha))))


def ppr_cheb_steps_for_tol(alpha, tol):
    """Steps of the Chebyshev-accelerated iteration for error ``tol``: 2 s^K / (1 + s^2K) <= tol with
    s = (1 - sqrt(1 - rho^2)) / rho, rho = 1 - alpha (plus a margin of 4 steps: fp32 round-off floor)."""
    import math
    rho = 1.0 - alpha
    s = (1.0 - math.sqrt(1.0 - rho * rho)) / rho
    k = 1
    while 2.0 * s ** k / (1.0 + s ** (2 * k)) > tol:
        k += 1
    return k + 4


def ppr_dense(ahat: NormalizedCSR, alpha, K=None, tol=1e-7, method="power"):
    """helpers.py:68-71 ``compute_ppr``: Pi = alpha (I - (1-alpha) A_hat)^-1 as a dense fp32 n x n
    CUDA tensor, by iteration on all n right-hand sides (csrc/ppr_dense.cu): ``method="power"`` is the
    plain fixed point (error (1-alpha)^K), ``"chebyshev"`` its Chebyshev acceleration (~4x fewer steps at
    alpha = 0.1, same bytes per step)."""
    lib = _lib.load()
    if ahat.val32 is None:
        raise ValueError("ppr_dense needs the fp32 values of A_hat")
    if method not in ("power", "chebyshev"):
        raise ValueError(f"unknown method {method!r}")
    n = ahat.n
    if method == "chebyshev":
        K = ppr_cheb_steps_for_tol(alpha, tol) if K is None else int(K)
        if K < 1:
            raise ValueError("the Chebyshev iteration needs K >= 1")
        dev = ahat.indices.device
        Pi = torch.empty((n, n), dtype=torch.float32, device=dev)
        scratch = torch.empty((n, n), dtype=torch.float32, device=dev) if K > 1 else None
        with torch.cuda.device(dev):
            rc = lib.ppnp_ppr_dense_cheb(_lib.ptr(ahat.indptr), _lib.ptr(ahat.indices), _lib.ptr(ahat.val32), n,
                                         float(alpha), K, _lib.ptr(Pi), _lib.ptr(scratch), _lib.current_stream())
        _lib.check(rc, "ppnp_ppr_dense_cheb")
        return Pi
    K = ppr_steps_for_tol(alpha, tol) if K is None else int(K)
    dev = ahat.indices.device
    Pi = torch.empty((n, n), dtype=torch.float32, device=dev)
    scratch = torch.empty((n, n), dtype=torch.float32, device=dev) if K > 1 else None
    with torch.cuda.device(dev):
        rc = lib.ppnp_ppr_dense(_lib.ptr(ahat.indptr), _lib.ptr(ahat.indices), _lib.ptr(ahat.val32), n,
                                float(alpha), K, _lib.ptr(Pi), _lib.ptr(scratch), _lib.current_stream())
    _lib.check(rc, "ppnp_ppr_dense")
    return Pi


def gather_gemm(Pi, H, idx=None, transpose=False, n_out=None):
    """model.py:63 ``ppr[idx] @ H`` / model.py:65 ``ppr @ H`` in fp32 (csrc/gather_gemm.cu).
    ``transpose=True`` gives the autograd adjoint ``ppr[idx].T @ H``."""
    lib = _lib.load()
    _require_cuda(Pi, H, idx)
    if Pi.dtype != torch.float32 or H.dtype != torch.float32:
        raise ValueError("gather_gemm is the fp32 path; use gather_gemm_bf16 for bf16 Pi")
    if Pi.stride(1) != 1:
        Pi = Pi.contiguous()
    H = H.contiguous()
    n = Pi.shape[1]
    if idx is not None:
        idx = idx.to(device=Pi.device, dtype=torch.int64).contiguous()
        m = idx.numel()
    else:
        m = Pi.shape[0]
    C = H.shape[1]
    if not transpose:
        if H.shape[0] != n:
            raise ValueError(f"H has {H.shape[0]} rows, Pi has {n} columns")
        out = torch.empty((m, C), dtype=torch.float32, device=Pi.device)
    else:
        if H.shape[0] != m:
            raise ValueError(f"H has {H.shape[0]} rows, expected {m}")
        out = torch.empty((n, C), dtype=torch.float32, device=Pi.device)
    if m == 0:
        return out.zero_()
    with torch.cuda.device(Pi.device):
        rc = lib.ppnp_gather_gemm_f32(_lib.ptr(Pi), Pi.stride(0), _lib.ptr(idx), m, n, _lib.ptr(H), H.stride(0), C,
                                      _lib.ptr(out), out.stride(0), int(bool(transpose)), _lib.current_stream())
    _lib.check(rc, "ppnp_gather_gemm_f32")
    return out


def to_bf16(x):
    """fp32 -> bf16 copy (round to nearest even) through the library's own kernel."""
    lib = _lib.load()
    _require_cuda(x)
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.ppnp_f32_to_bf16(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.current_stream())
    _lib.check(rc, "ppnp_f32_to_bf16")
    return out


def to_bf16_padded(Pi):
    """bf16 shadow of a dense fp32 matrix with the row stride padded to a multiple of 64 elements
    (16-byte aligned rows, whole k-blocks): what ``gather_gemm_bf16`` consumes.  Returns the
    n_rows x n_cols view into the padded buffer."""
    n_rows, n_cols = Pi.shape
    ld = ((n_cols + 63) // 64) * 64
    buf = torch.zeros((n_rows, ld), dtype=torch.bfloat16, device=Pi.device)
    # convert in row slabs of <= 256 MB so the fp32 source never needs a second full-size copy
    rows_per = max(1, (64 << 20) // max(n_cols, 1))
    for r0 in range(0, n_rows, rows_per):
        r1 = min(n_rows, r0 + rows_per)
        buf[r0:r1, :n_cols] = to_bf16(Pi[r0:r1])
    return buf[:, :n_cols]


def gather_gemm_bf16(Pi_bf16, H, idx=None):
    """The bf16 tensor-core form of model.py:63/65 (tcgen05.mma, fp32 accumulation in TMEM)."""
    lib = _lib.load()
    _require_cuda(Pi_bf16, H, idx)
    if Pi_bf16.dtype != torch.bfloat16:
        raise ValueError("Pi must be bfloat16")
    H = H.to(torch.float32).contiguous()
    n = Pi_bf16.shape[1]
    if idx is not None:
        idx = idx.to(device=Pi_bf16.device, dtype=torch.int64).contiguous()
        m = idx.numel()
    else:
        m = Pi_bf16.shape[0]
    C = H.shape[1]
    out = torch.empty((m, C), dtype=torch.float32, device=H.device)
    ws_bytes = lib.ppnp_gather_gemm_bf16_workspace_bytes(m, n, C)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=H.device)
    with torch.cuda.device(H.device):
        rc = lib.ppnp_gather_gemm_bf16(_lib.ptr(Pi_bf16), Pi_bf16.stride(0), _lib.ptr(idx), m, n, _lib.ptr(H),
                                       H.stride(0), C, _lib.ptr(out), out.stride(0), _lib.ptr(ws), ws_bytes,
                                       _lib.current_stream())
    _lib.check(rc, "ppnp_gather_gemm_bf16")
    return out


class _GatherGemmFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, Pi, idx):
        ctx.Pi, ctx.idx = Pi, idx
        return gather_gemm(Pi, H, idx)

    @staticmethod
    def backward(ctx, G):
        return gather_gemm(ctx.Pi, G.contiguous(), ctx.idx, transpose=True), None, None


def ppr_matmul(Pi, H, idx=None):
    """Differentiable (w.r.t. H) ``Pi[idx] @ H``; Pi is a constant buffer as in model.py:54."""
    return _GatherGemmFunction.apply(H, Pi, idx)


# --------------------------------------------------------------------------- batch-main path
def topk_thresh(ppr, k):
    """batch-main.py:115: k-th largest of every row (== ``ppr.topk(k, -1).values[:, -1]``)."""
    lib = _lib.load()
    _require_cuda(ppr)
    if ppr.stride(1) != 1:
        ppr = ppr.contiguous()
    n_rows, n_cols = ppr.shape
    th = torch.empty(n_rows, dtype=torch.float32, device=ppr.device)
    with torch.cuda.device(ppr.device):
        rc = lib.ppnp_topk_thresh(_lib.ptr(ppr), n_rows, n_cols, ppr.stride(0), int(k), _lib.ptr(th), _lib.current_stream())
    _lib.check(rc, "ppnp_topk_thresh")
    return th


def topk_sparsify_(ppr, k):
    """batch-main.py:115-116 in place on the dense CUDA matrix (column-broadcast quirk kept)."""
    lib = _lib.load()
    if ppr.stride(1) != 1:
        raise ValueError("ppr must be row-major")
    if ppr.shape[0] != ppr.shape[1]:
        raise ValueError("thresh[:, -1] only broadcasts against a square matrix (as in the reference)")
    th = topk_thresh(ppr, k)
    with torch.cuda.device(ppr.device):
        rc = lib.ppnp_topk_mask(_lib.ptr(ppr), ppr.shape[0], ppr.shape[1], ppr.stride(0), _lib.ptr(th), _lib.current_stream())
    _lib.check(rc, "ppnp_topk_mask")
    return th


@dataclass
class SparsePPR:
    """Compact CSR of the entries > 0 of a (sparsified) dense PPR matrix."""
    n_rows: int
    n_cols: int
    indptr: torch.Tensor    # int64 [n_rows + 1]
    indices: torch.Tensor   # int32 [nnz]
    val: torch.Tensor       # fp32 [nnz]


def dense_to_sparse_ppr(ppr):
    lib = _lib.load()
    _require_cuda(ppr)
    if ppr.stride(1) != 1:
        ppr = ppr.contiguous()
    n_rows, n_cols = ppr.shape
    row_nnz = torch.empty(n_rows, dtype=torch.int32, device=ppr.device)
    with torch.cuda.device(ppr.device):
        rc = lib.ppnp_dense_row_nnz(_lib.ptr(ppr), n_rows, n_cols, ppr.stride(0), _lib.ptr(row_nnz), _lib.current_stream())
        _lib.check(rc, "ppnp_dense_row_nnz")
        indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=ppr.device)
        indptr[1:] = torch.cumsum(row_nnz.to(torch.int64), 0)
        nnz = int(indptr[-1].item())
        indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=ppr.device)
        val = torch.empty(max(nnz, 1), dtype=torch.float32, device=ppr.device)
        rc = lib.ppnp_dense_to_csr(_lib.ptr(ppr), n_rows, n_cols, ppr.stride(0), _lib.ptr(indptr), _lib.ptr(indices),
                                   _lib.ptr(val), _lib.current_stream())
        _lib.check(rc, "ppnp_dense_to_csr")
    return SparsePPR(n_rows, n_cols, indptr, indices[:nnz], val[:nnz])


_COLMAP_MAX_COLS = 200 * 1024       # ppnp_batch_support_colmap: one byte of shared memory per column


def batch_support(sp_ppr: SparsePPR, idx_batch):
    """batch-main.py:140-141 on the compact matrix: bool mask ``sel`` over the n columns.  One launch also yields
    the column map line 142 needs (position of every column inside ``sel``, -1 outside); it rides on the returned
    tensor (``sel._ppnp_colmap``) so that ``batch_propagate`` with this very mask starts no further kernel."""
    lib = _lib.load()
    dev = sp_ppr.indices.device
    idx_batch = idx_batch.to(device=dev, dtype=torch.int64).contiguous()
    n = sp_ppr.n_cols
    if idx_batch.numel() and n <= _COLMAP_MAX_COLS:
        sel = torch.empty(n, dtype=torch.bool, device=dev)
        colmap = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.ppnp_batch_support_colmap(_lib.ptr(sp_ppr.indptr), _lib.ptr(sp_ppr.indices), _lib.ptr(idx_batch),
                                               idx_batch.numel(), n, _lib.ptr(sel), _lib.ptr(colmap), None, _lib.current_stream())
        _lib.check(rc, "ppnp_batch_support_colmap")
        sel._ppnp_colmap = (colmap, idx_batch)
        return sel
    mark = torch.zeros(n, dtype=torch.uint8, device=dev)
    if idx_batch.numel():
        with torch.cuda.device(mark.device):
            rc = lib.ppnp_batch_support(_lib.ptr(sp_ppr.indptr), _lib.ptr(sp_ppr.indices), _lib.ptr(idx_batch),
                                        idx_batch.numel(), _lib.ptr(mark), _lib.current_stream())
        _lib.check(rc, "ppnp_batch_support")
    return mark.to(torch.bool)


def _batch_propagate_raw(sp_ppr, idx_batch, colmap, Hsub, transpose, n_out_rows):
    lib = _lib.load()
    Hsub = Hsub.contiguous()
    C = Hsub.shape[1]
    B = idx_batch.numel()
    if not transpose:
        out = torch.empty((B, C), dtype=torch.float32, device=Hsub.device)
    else:
        out = torch.zeros((n_out_rows, C), dtype=torch.float32, device=Hsub.device)
    if B == 0:
        return out
    with torch.cuda.device(Hsub.device):
        rc = lib.ppnp_batch_propagate(_lib.ptr(sp_ppr.indptr), _lib.ptr(sp_ppr.indices), _lib.ptr(sp_ppr.val),
                                      _lib.ptr(idx_batch), B, _lib.ptr(colmap), _lib.ptr(Hsub), Hsub.stride(0), C,
                                      _lib.ptr(out), out.stride(0), int(bool(transpose)), _lib.current_stream())
    _lib.check(rc, "ppnp_batch_propagate")
    return out


class _BatchPropagateFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Hsub, sp_ppr, idx_batch, colmap):
        ctx.sp, ctx.idx, ctx.colmap, ctx.m = sp_ppr, idx_batch, colmap, Hsub.shape[0]
        return _batch_propagate_raw(sp_ppr, idx_batch, colmap, Hsub, False, None)

    @staticmethod
    def backward(ctx, G):
        return _batch_propagate_raw(ctx.sp, ctx.idx, ctx.colmap, G.contiguous(), True, ctx.m), None, None, None


def batch_propagate(sp_ppr: SparsePPR, idx_batch, sel, Hsub):
    """batch-main.py:142-146: ``ppr[idx_batch][:, sel] @ Hsub`` with Hsub = encoder(X[sel])
    (differentiable w.r.t. Hsub)."""
    dev = sp_ppr.indices.device
    idx_batch = idx_batch.to(device=dev, dtype=torch.int64).contiguous()
    cached = getattr(sel, "_ppnp_colmap", None)
    if cached is not None and (cached[1] is idx_batch or (cached[1].data_ptr() == idx_batch.data_ptr()
                                                          and cached[1].numel() == idx_batch.numel())):
        colmap = cached[0]                   # made by batch_support for this mask and this batch
    else:                                    # any other mask: positions inside it, -1 outside (arbitrary masks are valid)
        pos = torch.cumsum(sel.to(torch.int32), 0, dtype=torch.int32) - 1
        colmap = torch.where(sel, pos, torch.full_like(pos, -1)).contiguous()
    return _BatchPropagateFunction.apply(Hsub, sp_ppr, idx_batch, colmap)


# ------------------------------------------------------------------------------ partitioned helper
def gather_rows(src, idx, out):
    """out[i] = src[idx[i]] (fp32 rows); ``src`` may be a peer GPU's symmetric-memory view."""
    lib = _lib.load()
    _require_cuda(src, idx, out)
    n_rows = idx.numel()
    if n_rows == 0:
        return out
    if src.stride(1) != 1 or out.stride(1) != 1 or idx.dtype != torch.int64:
        raise ValueError("gather_rows needs row-major fp32 matrices and an int64 index")
    with torch.cuda.device(out.device):
        rc = lib.ppnp_gather_rows(_lib.ptr(src), src.stride(0), _lib.ptr(idx), n_rows, src.shape[1], _lib.ptr(out),
                                  out.stride(0), _lib.current_stream())
    _lib.check(rc, "ppnp_gather_rows")
    return out


# ------------------------------------------------------------------------------ graph standardisation
def graph_standardize(indptr, indices, make_unweighted=True, make_undirected=True, no_self_loops=True, select_lcc=True):
    """ppnp/data/sparsegraph.py:191-222 ``SparseGraph.standardize`` on the GPU (csrc/standardize.cu).

    indptr / indices: CSR pattern of the raw adjacency as CUDA tensors (weights are never read: the
    pipeline of main.py:75 sets them to 1 first).  Returns ``(indptr, indices, keep)``: the canonical
    CSR of the standardised graph (int32 when it fits, the contract of ``csr_normalize``) and the
    original ids of the kept nodes (int64, ascending) for subsetting attributes and labels.
    """
    if not make_unweighted:
        raise NotImplementedError("only the make_unweighted=True pipeline (main.py:75) exists on the GPU")
    lib = _lib.load()
    _require_cuda(indptr, indices)
    dev = indices.device
    indptr = indptr.to(torch.int64).contiguous()
    indices = indices.to(torch.int32).contiguous()
    n = indptr.numel() - 1
    nnz = indices.numel()
    if n <= 0:
        raise ValueError("empty graph")
    flags = ((_lib.STD_UNDIRECTED if make_undirected else 0) | (_lib.STD_NO_SELF_LOOPS if no_self_loops else 0) |
             (_lib.STD_LCC if select_lcc else 0))
    cap = max(1, 2 * nnz if make_undirected else nnz)
    out_indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    out_indices = torch.empty(cap, dtype=torch.int32, device=dev)
    out_keep = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.empty(3, dtype=torch.int64, device=dev)
    ws_bytes = lib.ppnp_graph_standardize_workspace_bytes(n, nnz, flags)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.ppnp_graph_standardize(_lib.ptr(indptr), _lib.ptr(indices), n, nnz, flags, _lib.ptr(out_indptr),
                                        _lib.ptr(out_indices), _lib.ptr(out_keep), _lib.ptr(counts), _lib.ptr(ws), ws_bytes,
                                        _lib.current_stream())
    _lib.check(rc, "ppnp_graph_standardize")
    n_keep, nnz_out, status = (int(x) for x in counts.tolist())
    if status != 0:
        raise ValueError("graph_standardize: a column index lies outside [0, n)")
    ip = out_indptr[:n_keep + 1]
    if nnz_out < (1 << 31):
        ip = ip.to(torch.int32)
    return ip, out_indices[:nnz_out], out_keep[:n_keep].to(torch.int64)


# ------------------------------------------------------------------------------ sparse-input first layer
class SparseInput:
    """The attribute matrix X (n x F_in, ~2 % dense; main.py:90-92 densifies it) kept sparse for the
    encoder's first layer (model.py:47-48: Dropout then CustomLinear), SURVEY.md section 8f rank 3.

    ``dropout(X) @ W`` is an SpMM whose "adjacency" is X's CSR and whose feature matrix is W (F_in x hidden):
    the propagation kernel runs it as it is (stored-value form, alpha = 0, accumulate epilogue onto a zeroed
    output).  The gradient ``dropout(X)^T @ dOut`` is the same kernel over the stream of X^T.  Dropout is
    applied to the stored values of the stream (one Bernoulli draw per stored entry, scaled by 1/(1-p));
    the dense n x F_in mask of nn.Dropout never exists.  Two stream plans are built once; rows (or
    columns) without entries are simply absent from their stream and stay zero."""

    def __init__(self, indptr, indices, values, n_cols, chunk_edges=256):
        dev = indices.device            # index bookkeeping runs wherever the arrays live; the launches need CUDA
        ip = indptr.to(torch.int64)
        self.n_rows, self.n_cols = int(ip.numel()) - 1, int(n_cols)
        idx = indices.to(torch.int64)
        nnz = int(idx.numel())
        self.nnz = nnz
        self.values = values.to(torch.float32).contiguous()
        cnt = ip[1:] - ip[:-1]
        row_of = torch.repeat_interleave(torch.arange(self.n_rows, device=dev), cnt)
        # X^T as CSR: entries sorted by (column, row)
        tperm = torch.sort(idx * self.n_rows + row_of, stable=True).indices
        tcnt = torch.bincount(idx, minlength=self.n_cols)
        tip = torch.zeros(self.n_cols + 1, dtype=torch.int64, device=dev)
        tip[1:] = torch.cumsum(tcnt, 0)
        self.fwd, self.fwd_edge = self._stream(ip, idx.to(torch.int32), cnt, chunk_edges, None)
        self.bwd, self.bwd_edge = self._stream(tip, row_of[tperm].to(torch.int32), tcnt, chunk_edges, tperm)

    @staticmethod
    def _stream(ip, cols, cnt, chunk_edges, edge_of_csr):
        """Stream plan over the non-empty rows (degree order) + the original entry behind every stream position."""
        dev = cols.device
        rows = torch.nonzero(cnt > 0).flatten()
        order = rows[torch.sort(cnt[rows], descending=True, stable=True).indices]
        if order.numel() == 0:
            return None, None
        plan = build_stream_plan(ip, cols, torch.ones(cols.numel(), dtype=torch.float32, device=dev), chunk_edges, order, subset=True)
        L = cnt[order]
        a = torch.cumsum(L, 0) - L
        src = torch.repeat_interleave(ip[:-1][order] - a, L) + torch.arange(int(L.sum()), device=dev, dtype=torch.int64)
        if edge_of_csr is not None:
            src = edge_of_csr[src]
        return plan, src

    @classmethod
    def from_dense(cls, X, chunk_edges=256):
        """From the dense tensor main.py:91-92 builds (torch's dense -> CSR conversion is plumbing)."""
        s = X.to_sparse_csr()
        return cls(s.crow_indices(), s.col_indices(), s.values(), X.shape[1], chunk_edges)

    def _with_values(self, plan, edge_of_pos, scale_per_entry):
        import dataclasses
        v = self.values if scale_per_entry is None else self.values * scale_per_entry
        vals = torch.zeros(plan.cols.numel(), dtype=torch.float32, device=v.device)
        vals[: edge_of_pos.numel()] = v[edge_of_pos]
        return dataclasses.replace(plan, vals=vals, _struct=None)

    def _run(self, plan, Zin, out):
        lib = _lib.load()
        _require_cuda(Zin, plan.cols)
        F = Zin.shape[1]
        partial = None
        if plan.n_slots:
            partial = torch.empty(plan.n_slots * F, dtype=torch.float32, device=Zin.device)
        with torch.cuda.device(Zin.device):
            rc = lib.ppnp_spmm_step(plan.struct(), _lib.ptr(Zin), _lib.ptr(out), _lib.ptr(out), _lib.ptr(partial), F, F,
                                    0.0, _lib.EPI_PLAIN | _lib.EPI_ACC, 1, _lib.current_stream())
        _lib.check(rc, "ppnp_spmm_step")
        return out

    def matmul(self, W, scale_per_entry=None):
        """(X * scale) @ W  ->  [n_rows, hidden]."""
        W = W.contiguous()
        out = torch.zeros((self.n_rows, W.shape[1]), dtype=torch.float32, device=W.device)
        if self.fwd is None:
            return out
        return self._run(self._with_values(self.fwd, self.fwd_edge, scale_per_entry), W, out)

    def rmatmul(self, G, scale_per_entry=None):
        """(X * scale)^T @ G  ->  [n_cols, hidden]."""
        G = G.contiguous()
        out = torch.zeros((self.n_cols, G.shape[1]), dtype=torch.float32, device=G.device)
        if self.bwd is None:
            return out
        return self._run(self._with_values(self.bwd, self.bwd_edge, scale_per_entry), G, out)


class _SparseLinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W, sx, scale):
        ctx.sx, ctx.scale = sx, scale
        return sx.matmul(W, scale)

    @staticmethod
    def backward(ctx, G):
        return ctx.sx.rmatmul(G, ctx.scale), None, None


def sparse_first_layer(sx: SparseInput, W, p=0.5, training=True, generator=None):
    """model.py:47-48 ``CustomLinear(Dropout(p)(X))`` (bias-free part) with X kept sparse: differentiable w.r.t. W."""
    if W.dtype != torch.float32 or W.dim() != 2 or W.shape[0] != sx.n_cols:
        raise ValueError(f"W must be a float32 [{sx.n_cols}, hidden] matrix")
    scale = None
    if training and p > 0.0:
        if p >= 1.0:
            scale = torch.zeros(sx.nnz, dtype=torch.float32, device=W.device)
        else:
            keep = torch.empty(sx.nnz, dtype=torch.float32, device=W.device).bernoulli_(1.0 - p, generator=generator)
            scale = keep / (1.0 - p)
    return _SparseLinearFunction.apply(W, sx, scale)
