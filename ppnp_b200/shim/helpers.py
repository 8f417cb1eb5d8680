"""Drop-in replacement for the reference's flat module ``helpers`` (helpers.py), importable by
bare name when this directory precedes the reference on sys.path:

    sys.path[:0] = [<repo>/ppnp_b200/shim]; runpy.run_path('<reference>/main.py', run_name='__main__')

main.py:31 / batch-main.py:32 import ``set_seeds, compute_ppr, SimpleEarlyStopping`` from here.

  calc_A_hat(adj, mode)            helpers.py:58-66  -> GPU kernels (csrc/csr_normalize.cu), returned
                                                        as the same scipy CSR (fp64, bit-exact)
  compute_ppr(adj, alpha, mode)    helpers.py:68-71  -> GPU power iteration (csrc/ppr_dense.cu), returned
                                                        as a host fp32 ndarray that
                                                        ``torch.FloatTensor(...)`` (main.py:106) takes
                                                        zero-copy
  set_seeds, SimpleEarlyStopping   helpers.py:13-55  -> training-harness utilities with the reference's
                                                        behaviour (they are not on the hot path)

Environment switches (main.py has no flag for them and stays unchanged):
  PPNP_MODE = exact | appnp   exact PPNP (default) or K-step APPNP; in appnp mode compute_ppr only
                              records the normalised graph for model.PPNP and returns a placeholder
  PPNP_K                      APPNP steps (default 10)
  PPNP_PPR_TOL                stop tolerance of the power iteration building Pi (default 1e-7)
  PPNP_PPR_METHOD = chebyshev | power   Chebyshev-accelerated iteration (default: 40 steps instead of 153 at
                              alpha = 0.1, measured 75 ms against 220 ms at PubMed shape, 9e-8 apart) or the plain fixed point
"""
import os
import random

import numpy as np
import scipy.sparse as sp
import torch

import ppnp_b200 as _P

# module-level hand-over to model.PPNP (main.py passes only Pi to the model, never the adjacency)
LAST_GRAPH = {"ahat": None, "alpha": None, "mode": None}


def set_seeds(seed):
    """helpers.py:13-17: python / numpy / torch / cuda generators at seed+1 .. seed+4."""
    random.seed(seed + 1)
    np.random.seed(seed + 2)
    torch.manual_seed(seed + 3)
    torch.cuda.manual_seed(seed + 4)


class SimpleEarlyStopping:
    """helpers.py:19-55: patience counter over (accuracy, -loss); the counter drops only when BOTH
    are worse than their running best, stops at zero, and any other epoch resets it; the best epoch
    is the lexicographic maximum of (acc, -loss) and keeps the caller's ``record``."""

    def __init__(self, model, patience=100, store_weights=False):
        self.model = model
        self.patience = patience
        self.max_patience = patience
        self.store_weights = store_weights
        self.record = (None,)
        self.best_acc = -np.inf
        self.best_nloss = -np.inf
        self.best_epoch = -1
        self.best_epoch_score = (-np.inf, -np.inf)

    def should_stop(self, acc, loss, epoch, record=None):
        nloss = -loss
        worse_everywhere = acc < self.best_acc and nloss < self.best_nloss
        if worse_everywhere:
            self.patience -= 1
            return self.patience == 0
        self.patience = self.max_patience
        self.best_acc = max(self.best_acc, acc)
        self.best_nloss = max(self.best_nloss, nloss)
        score = (acc, nloss)
        if score > self.best_epoch_score:
            self.best_epoch, self.best_epoch_score = epoch, score
            if self.store_weights:
                self.best_state = {name: t.cpu() for name, t in self.model.state_dict().items()}
            if record:
                self.record = record
        return False


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("ppnp_b200: no CUDA device -- the propagation path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _normalize_on_gpu(adj, mode, want_val64):
    adj = sp.csr_matrix(adj)
    if not adj.has_sorted_indices:
        adj = adj.sorted_indices()
    dev = _device()
    indptr = torch.from_numpy(np.ascontiguousarray(adj.indptr, dtype=np.int32)).to(dev)
    indices = torch.from_numpy(np.ascontiguousarray(adj.indices, dtype=np.int32)).to(dev)
    data = None
    if adj.nnz and not np.all(adj.data == 1):
        data = torch.from_numpy(np.ascontiguousarray(adj.data, dtype=np.float32)).to(dev)
    return _P.csr_normalize(indptr, indices, data, mode, want_val64=want_val64)


def calc_A_hat(adj, mode):
    """helpers.py:58-66 on the GPU; returns scipy CSR fp64 with the reference's exact structure and
    values (``None`` for an unknown mode, like the reference's fall-through)."""
    if mode not in ("sym", "rw"):
        return None
    ahat = _normalize_on_gpu(adj, mode, want_val64=True)
    n = ahat.n
    return sp.csr_matrix((ahat.val64.cpu().numpy(), ahat.indices.cpu().numpy(), ahat.indptr.cpu().numpy()), shape=(n, n))


def compute_ppr(adj, alpha, mode="sym"):
    """helpers.py:68-71.  Host float32 [n, n] array of alpha (I - (1-alpha) A_hat)^-1 (pinned memory
    behind it, so the caller's ``.cuda()`` is a plain DMA)."""
    ahat = _normalize_on_gpu(adj, mode, want_val64=False)
    LAST_GRAPH.update(ahat=ahat, alpha=float(alpha), mode=mode)
    if os.environ.get("PPNP_MODE", "exact").lower() == "appnp":
        return np.zeros((1, 1), dtype=np.float32)          # placeholder; model.PPNP propagates on the graph
    tol = float(os.environ.get("PPNP_PPR_TOL", "1e-7"))
    Pi = _P.ppr_dense(ahat, float(alpha), tol=tol, method=os.environ.get("PPNP_PPR_METHOD", "chebyshev"))
    host = torch.empty(Pi.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(Pi)
    torch.cuda.synchronize()
    return host.numpy()
