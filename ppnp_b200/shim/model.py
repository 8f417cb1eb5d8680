"""Drop-in replacement for the reference's flat module ``model`` (model.py): ``from model import PPNP``
(main.py:30, batch-main.py:31) resolves here when this directory precedes the reference on sys.path.

Same constructor, same ``encoder`` (so ``torch.manual_seed`` at main.py:104 yields the reference's
initial weights), same buffer name ``ppr``, same three-way ``forward`` -- but the propagation runs
in ppnp_b200's CUDA kernels:

  forward(X, idx)        model.py:63  ppr[idx] @ H   -> gather-GEMM reading only the |idx| rows
                                                        (csrc/gather_gemm.cu fp32; gather_gemm_tc.cu when
                                                        PPNP_GEMM=bf16), or, with PPNP_MODE=appnp, K steps of
                                                        Z <- (1-a) A_hat Z + a H (csrc/appnp_spmm.cu), then [idx]
  forward(X, ppr=sub)    model.py:65  ppr @ H        -> the same GEMM on a dense ``sub``; on the compact
                                                        top-k form when ``sub`` came from ``model.ppr[idx_batch]``
                                                        (batch-main.py:140-146, csrc/batch.cu)
  neither                model.py:67  raise Exception()

The MLP feature transform (Dropout, Linear, ReLU, Dropout, Linear) stays in PyTorch (north_star).
"""
import math
import os

import torch
from torch import nn

import ppnp_b200 as _P


class CustomLinear(nn.Module):
    """model.py:13-38: weight stored [in, out], kaiming-uniform with fan_out, ``input @ weight``."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.Tensor(in_features, out_features))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, mode="fan_out", a=math.sqrt(5))
        if self.bias is not None:
            _, fan_out = nn.init._calculate_fan_in_and_fan_out(self.weight)
            nn.init.uniform_(self.bias, -1 / math.sqrt(fan_out), 1 / math.sqrt(fan_out))

    def forward(self, x):
        return x @ self.weight if self.bias is None else torch.addmm(self.bias, x, self.weight)


# ---------------------------------------------------------------- batch-main.py:140-142 on the compact form
class _BatchRows:
    """``model.ppr[idx_batch]`` (batch-main.py:140), not materialised."""

    def __init__(self, owner, idx):
        self.owner, self.idx = owner, idx.to(owner.device)
        self.support = None      # set by (rows > 0).any(dim=0)

    def dense(self):
        return torch.Tensor.__getitem__(self.owner.as_subclass(torch.Tensor), self.idx)

    def __gt__(self, other):
        if other == 0:
            return _BatchRowsPositive(self)
        return self.dense() > other

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and key[0] == slice(None) and torch.is_tensor(key[1]) \
                and key[1].dtype == torch.bool and key[1] is self.support:
            # batch-main.py:142 with the mask of :141.  The compact kernels index Hsub by the position of a column
            # inside THIS mask (colmap = cumsum(sel) - 1, unchecked): any other mask takes the dense path below,
            # which is valid for arbitrary masks like the reference's ppr_sub[:, mask].
            return _BatchSub(self, key[1])
        return self.dense()[key]

    def __getattr__(self, name):            # anything else behaves like the dense B x n tensor
        return getattr(self.dense(), name)


class _BatchRowsPositive:
    """``(ppr_sub > 0)`` (batch-main.py:141): only ``.any(dim=0)`` is ever asked of it."""

    def __init__(self, rows):
        self.rows = rows

    def any(self, dim=None):
        if dim in (0, -2):
            sel = _P.batch_support(self.rows.owner.compact(), self.rows.idx)
            self.rows.support = sel          # the one mask the compact product below is valid for
            return sel
        return (self.rows.dense() > 0).any(dim=dim)

    def __getattr__(self, name):
        return getattr(self.rows.dense() > 0, name)


class _BatchSub:
    """``ppr_sub[:, sel]`` (batch-main.py:142): consumed by ``PPNP.forward(X_batch, ppr=...)``."""

    def __init__(self, rows, sel):
        self.rows, self.sel = rows, sel

    def dense(self):
        return self.rows.dense()[:, self.sel]

    def __getattr__(self, name):
        return getattr(self.dense(), name)


class TopkPPR(torch.Tensor):
    """The ``ppr`` buffer: an ordinary dense tensor (state_dict, ``.cuda()``, eval through
    model.py:63 all see plain data) whose row gather by a 1-D index tensor on a CUDA device returns
    the lazy batch view above, backed by a compact CSR built once from the entries > 0."""

    @staticmethod
    def __new__(cls, data):
        return torch.Tensor._make_subclass(cls, data, False)

    def compact(self):
        key = (self.data_ptr(), self._version)
        cached = getattr(self, "_compact", None)
        if cached is None or cached[0] != key:
            cached = (key, _P.dense_to_sparse_ppr(self.as_subclass(torch.Tensor)))
            self._compact = cached
        return cached[1]

    def __getitem__(self, key):
        if torch.is_tensor(key) and key.dim() == 1 and key.dtype == torch.int64 and self.is_cuda and self.dim() == 2:
            return _BatchRows(self, key)
        return torch.Tensor.__getitem__(self.as_subclass(torch.Tensor), key)


class PPNP(nn.Module):
    """model.py:41-67 with the propagation in CUDA."""

    def __init__(self, n_features, n_classes, ppr, hidden_dim=64, drop_prob=0.5, bias=False):
        super().__init__()
        n_classes = int(n_classes)          # main.py:107 passes a 0-d LongTensor
        self.encoder = nn.Sequential(
            nn.Dropout(drop_prob),
            CustomLinear(n_features, hidden_dim, bias=bias),
            nn.ReLU(inplace=True),
            nn.Dropout(drop_prob),
            nn.Linear(hidden_dim, n_classes, bias=bias),
        )
        self.mode = os.environ.get("PPNP_MODE", "exact").lower()
        self.gemm = os.environ.get("PPNP_GEMM", "fp32").lower()
        self.K = int(os.environ.get("PPNP_K", "10"))
        sparse_batches = os.environ.get("PPNP_BATCH", "sparse").lower() == "sparse"
        self.register_buffer("ppr", TopkPPR(ppr) if (sparse_batches and self.mode == "exact") else ppr)
        self._reg_params = list(self.encoder[1].parameters())
        self._graph = None
        self._alpha = None
        self._ppr_bf16 = None
        # PPNP_SPARSE_X=1: the first encoder layer reads X as a sparse matrix (ppnp_b200.SparseInput) in the
        # full-batch branch (model.py:63); main.py still hands over the dense tensor, its CSR is built once
        self._sparse_x = os.environ.get("PPNP_SPARSE_X", "0") == "1"
        self._sx, self._sx_key = None, None
        # PPNP_FUSED_TAIL=1 (APPNP mode): the encoder's last linear (model.py:51) writes D^-1/2 H straight from the hidden
        # activations and the K steps run value-free from the first one (ppnp_b200.appnp_fused_tail): H is never
        # materialised and the stored values of A_hat are never read
        self._fused_tail = os.environ.get("PPNP_FUSED_TAIL", "0") == "1"
        if self.mode == "appnp":
            import helpers as _h            # the shim module that recorded the graph (helpers.compute_ppr)
            if _h.LAST_GRAPH["ahat"] is None:
                raise RuntimeError("PPNP_MODE=appnp: call helpers.compute_ppr(adj, alpha) before building PPNP")
            self._graph = _P.PropagationGraph(_h.LAST_GRAPH["ahat"])
            self._alpha = _h.LAST_GRAPH["alpha"]

    def get_norm(self):
        return sum(torch.sum(p ** 2) for p in self._reg_params)

    def _apply_ppr(self, ppr, H, idx):
        if self.gemm == "bf16":
            if ppr is self.ppr:
                if self._ppr_bf16 is None or self._ppr_bf16.device != H.device:
                    self._ppr_bf16 = _P.to_bf16_padded(self.ppr.as_subclass(torch.Tensor))
                shadow = self._ppr_bf16
            else:
                shadow = _P.to_bf16_padded(ppr)
            return _GemmBf16.apply(H, shadow, ppr.as_subclass(torch.Tensor), idx)
        return _P.ppr_matmul(ppr.as_subclass(torch.Tensor), H, idx)

    def _encode_full(self, X):
        """model.py:46-52 on the whole attribute matrix; with PPNP_SPARSE_X=1 the Dropout + CustomLinear pair
        (model.py:47-48) runs over X's stored entries only."""
        if not (self._sparse_x and X.is_cuda and X.dim() == 2):
            return self.encoder(X)
        key = (X.data_ptr(), tuple(X.shape), X._version)
        if self._sx_key != key:
            self._sx, self._sx_key = _P.SparseInput.from_dense(X), key
        drop, lin = self.encoder[0], self.encoder[1]
        h = _P.sparse_first_layer(self._sx, lin.weight, drop.p, self.training)
        if lin.bias is not None:
            h = h + lin.bias
        return self.encoder[4](self.encoder[3](self.encoder[2](h)))

    def forward(self, X, idx=None, ppr=None):
        if idx is not None:
            if self.mode == "appnp" and self._fused_tail and not self._sparse_x and self._graph.unit_weights:
                A1 = self.encoder[3](self.encoder[2](self.encoder[1](self.encoder[0](X))))       # model.py:47-50
                last = self.encoder[4]
                return _P.appnp_fused_tail(A1, last.weight, last.bias, self._graph, self.K, self._alpha)[idx]
            H = self._encode_full(X)
            if self.mode == "appnp":
                return _P.appnp(H, self._graph, self.K, self._alpha)[idx]
            return self._apply_ppr(self.ppr, H, idx)
        elif ppr is not None:
            H = self.encoder(X)
            if isinstance(ppr, _BatchSub):
                return _P.batch_propagate(ppr.rows.owner.compact(), ppr.rows.idx, ppr.sel, H)
            return self._apply_ppr(ppr, H, None)
        else:
            raise Exception()


class _GemmBf16(torch.autograd.Function):
    """bf16 tensor-core forward (1e-2 path); the backward uses the fp32 adjoint kernel."""

    @staticmethod
    def forward(ctx, H, shadow, Pi32, idx):
        ctx.Pi32, ctx.idx = Pi32, idx
        return _P.gather_gemm_bf16(shadow, H, idx)

    @staticmethod
    def backward(ctx, G):
        return _P.gather_gemm(ctx.Pi32, G.contiguous(), ctx.idx, transpose=True), None, None, None
