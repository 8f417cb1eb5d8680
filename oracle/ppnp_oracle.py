"""TEST INFRASTRUCTURE -- CPU restatement (numpy / scipy) of the PPNP/APPNP propagation path of
bkj/ppnp.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under ppnp_b200/ does.

Every function cites the reference lines it restates (paths into /root/reference).  Parity
status: PINNED -- tests/test_oracle_golden.py checks each function against golden vectors that
oracle/gen_golden.py produced by importing and running the reference itself (helpers.calc_A_hat,
helpers.compute_ppr, model.PPNP.forward, the literal batch-main.py lines) on the two graphs the
reference ships (Cora-ML, CiteSeer).  APPNP's K-step iteration does not exist in the reference;
it is pinned through its K -> inf limit against compute_ppr (KAT-1), H = I (KAT-2) and
adjointness (KAT-3), SURVEY.md section 8(c).
"""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- helpers.py
def calc_A_hat(adj, mode="sym"):
    """helpers.py:58-66.  A = adj + I (59); D = row sums (60); 'sym': D^-1/2 A D^-1/2 (61-63);
    'rw': D^-1 A (64-66).  Returns scipy CSR fp64 with sorted indices."""
    n = adj.shape[0]
    A = (adj + sp.eye(n)).tocsr()
    A.sort_indices()
    D = np.asarray(A.sum(axis=1)).ravel()
    if mode == "sym":
        d = 1.0 / np.sqrt(D)
        rows = np.repeat(np.arange(n), np.diff(A.indptr))
        data = (d[rows] * A.data) * d[A.indices]          # (D_i^-1/2 a_ij) D_j^-1/2, that order
    elif mode == "rw":
        d = 1.0 / D
        rows = np.repeat(np.arange(n), np.diff(A.indptr))
        data = d[rows] * A.data
    else:
        raise ValueError(mode)
    return sp.csr_matrix((data, A.indices.copy(), A.indptr.copy()), shape=(n, n))


def compute_ppr(adj, alpha, mode="sym"):
    """helpers.py:68-71.  alpha * inv(I - (1-alpha) A_hat), dense fp64."""
    A_hat = calc_A_hat(adj, mode)
    A_inner = sp.eye(adj.shape[0]) - (1 - alpha) * A_hat
    return alpha * np.linalg.inv(A_inner.toarray())


# ----------------------------------------------------------------------------- model.py
def ppnp_forward(ppr, H, idx=None):
    """model.py:61-65 with H = encoder(X) already evaluated: ppr[idx] @ H (63) or ppr @ H (65)."""
    if idx is not None:
        return ppr[idx] @ H
    return ppr @ H


def ppnp_forward_grad(ppr, G, idx=None):
    """Autograd of model.py:63/65 w.r.t. H: dH = ppr[idx]^T @ dlogits (no gradient to the buffer)."""
    if idx is not None:
        return ppr[idx].T @ G
    return ppr.T @ G


# ------------------------------------------------------------------------ APPNP (north_star)
def appnp(A_hat, H, alpha, K, dtype=np.float64):
    """Z_0 = H; Z_{k+1} = (1-alpha) A_hat Z_k + alpha H  (BASELINE.json north_star; A_hat from
    helpers.py:58-63).  Not in the reference -- see the module docstring for how it is pinned."""
    A = A_hat.astype(dtype)
    H = np.asarray(H, dtype=dtype)
    Z = H.copy()
    for _ in range(K):
        Z = (1 - alpha) * (A @ Z) + alpha * H
    return Z


# --------------------------------------------------------------------------- batch-main.py
def topk_thresh(ppr, k):
    """batch-main.py:115: thresh, _ = ppr.topk(k, axis=-1); thresh[:, -1] (k-th largest per row)."""
    part = np.partition(ppr, ppr.shape[1] - k, axis=1)
    return part[:, ppr.shape[1] - k].copy()


def topk_sparsify(ppr, k):
    """batch-main.py:115-116 (out of place).  ``ppr[ppr < thresh[:, -1]] = 0``: thresh[:, -1] has
    shape [n] and broadcasts along the last axis, i.e. entry (i, j) is compared with the k-th
    largest of ROW j."""
    th = topk_thresh(ppr, k)
    out = ppr.copy()
    out[out < th[None, :]] = 0
    return out


def batch_step(ppr_sparsified, idx_batch, H_full):
    """batch-main.py:140-146 with the encoder output given for all rows:
    ppr_sub = ppr[idx_batch] (140); sel = (ppr_sub > 0).any(0) (141); ppr_sub[:, sel] (142);
    logits = ppr_sub @ encoder(X[sel]) (144-146, model.py:65).  Returns (logits, sel)."""
    ppr_sub = ppr_sparsified[idx_batch]
    sel = (ppr_sub > 0).any(axis=0)
    ppr_sub = ppr_sub[:, sel]
    return ppr_sub @ H_full[sel], sel


# ------------------------------------------------------------------- C restatement (ppnp_oracle.c)
_clib = None


def standardize(indptr, indices, data=None, make_unweighted=True, make_undirected=True, no_self_loops=True,
                select_lcc=True):
    """ppnp/data/sparsegraph.py:191-222 ``SparseGraph.standardize`` restated on CSR arrays in numpy.

    to_unweighted (:150-154): every stored entry becomes 1.  to_undirected (:127-148): A + A.T with
    the doubly-present entries (opposing pairs, self loops) counted once -- for unit weights the
    pattern union with all values 1.  remove_self_loops (:381-395): diagonal entries dropped.
    largest_connected_components (:355-379): scipy's weakly connected components (numbered by their
    smallest node), ``np.argsort(sizes)[::-1][:1]`` picks the one to keep, create_subgraph (:300-352)
    keeps its nodes in ascending order and relabels.  Returns (indptr int64, indices int32, kept node
    ids int64); the data of the result is all ones.  Only the unit-weight pipeline is restated
    (``make_unweighted=True``, what main.py:75 runs); stored explicit zeros are not supported.
    """
    if not make_unweighted:
        raise NotImplementedError("only the make_unweighted=True pipeline of main.py:75 is restated")
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    if data is not None and (np.asarray(data) == 0).any():
        raise ValueError("explicit zeros in the adjacency are not supported")
    n = len(indptr) - 1
    r = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    c = indices
    if make_undirected:
        r, c = np.concatenate([r, c]), np.concatenate([c, r])
    if no_self_loops:
        m = r != c
        r, c = r[m], c[m]
    key = np.unique(r * n + c)
    r, c = key // n, key % n
    keep = np.arange(n, dtype=np.int64)
    if select_lcc and n > 0:
        lab = np.arange(n, dtype=np.int64)          # -> smallest node id of the (weak) component
        while True:
            new = lab.copy()
            m = np.minimum(lab[r], lab[c])
            np.minimum.at(new, r, m)
            np.minimum.at(new, c, m)
            new = new[new]
            if np.array_equal(new, lab):
                break
            lab = new
        roots, sizes = np.unique(lab, return_counts=True)   # roots ascending == scipy's component numbering
        best = roots[np.argsort(sizes)[::-1][:1]]             # the literal expression of sparsegraph.py:374
        keep = np.nonzero(np.isin(lab, best))[0].astype(np.int64)
        newid = np.full(n, -1, dtype=np.int64)
        newid[keep] = np.arange(len(keep))
        sel = newid[r] >= 0
        r, c = newid[r[sel]], newid[c[sel]]
    out_indptr = np.zeros(len(keep) + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=len(keep)), out=out_indptr[1:])
    return out_indptr, c.astype(np.int32), keep


def clib():
    """Load oracle/_build/libppnp_oracle.so (built by oracle/Makefile or __graft_entry__.build)."""
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "_build", "libppnp_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle`")
        lib = C.CDLL(path)
        p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        lib.oracle_num_threads.restype = C.c_int
        lib.oracle_rmat_edges.restype = i64
        lib.oracle_rmat_edges.argtypes = [C.c_uint64, i32, i64, i64, i64, p, p]
        lib.oracle_sym_csr.restype = i64
        lib.oracle_sym_csr.argtypes = [i64, i64, p, p, p, p]
        lib.oracle_a_hat.restype = i64
        lib.oracle_a_hat.argtypes = [i64, p, p, p, i32, p, p, p, p]
        lib.oracle_appnp_f64.restype = None
        lib.oracle_appnp_f64.argtypes = [i64, p, p, p, p, p, p, i64, i32, C.c_double]
        lib.oracle_appnp_f32.restype = None
        lib.oracle_appnp_f32.argtypes = [i64, p, p, p, p, p, p, i64, i32, C.c_float]
        lib.oracle_topk_thresh.restype = None
        lib.oracle_topk_thresh.argtypes = [i64, i64, p, i32, p]
        lib.oracle_topk_mask.restype = None
        lib.oracle_topk_mask.argtypes = [i64, p, p]
        _clib = lib
    return _clib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def rmat_graph(n, raw_draws, scale, seed=0):
    """SURVEY.md 8(d) recipe: R-MAT (0.57, 0.19, 0.19, 0.05) draws, ids >= n and loops dropped,
    symmetrised, de-duplicated -> canonical CSR (indptr int64, indices int32) of adj."""
    lib = clib()
    src = np.empty(raw_draws, dtype=np.int32)
    dst = np.empty(raw_draws, dtype=np.int32)
    m = lib.oracle_rmat_edges(seed, scale, n, 0, raw_draws, _ptr(src), _ptr(dst))
    indptr = np.empty(n + 1, dtype=np.int64)
    indices = np.empty(2 * m, dtype=np.int32)
    nnz = lib.oracle_sym_csr(n, m, _ptr(src), _ptr(dst), _ptr(indptr), _ptr(indices))
    return indptr, indices[:nnz].copy()


def c_a_hat(indptr, indices, data=None, mode="sym"):
    """ppnp_oracle.c oracle_a_hat (helpers.py:58-66) -> (indptr int64, indices int32, val fp64, D fp64)."""
    lib = clib()
    n = len(indptr) - 1
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    if data is not None:
        data = np.ascontiguousarray(data, dtype=np.float32)
    oip = np.empty(n + 1, dtype=np.int64)
    oidx = np.empty(len(indices) + n, dtype=np.int32)
    oval = np.empty(len(indices) + n, dtype=np.float64)
    odeg = np.empty(n, dtype=np.float64)
    nnz = lib.oracle_a_hat(n, _ptr(indptr), _ptr(indices), _ptr(data) if data is not None else None,
                           0 if mode == "sym" else 1, _ptr(oip), _ptr(oidx), _ptr(oval), _ptr(odeg))
    return oip, oidx[:nnz].copy(), oval[:nnz].copy(), odeg


def c_appnp_f32(indptr, indices, val, H, K, alpha):
    """ppnp_oracle.c oracle_appnp_f32: the multi-threaded fp32 CPU port bench.py times."""
    lib = clib()
    n, F = H.shape
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    H = np.ascontiguousarray(H, dtype=np.float32)
    Z = np.empty_like(H)
    scratch = np.empty_like(H)
    lib.oracle_appnp_f32(n, _ptr(indptr), _ptr(indices), _ptr(val), _ptr(H), _ptr(Z), _ptr(scratch), F, K, alpha)
    return Z


def c_appnp_f64(indptr, indices, val, H, K, alpha):
    lib = clib()
    n, F = H.shape
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    H = np.ascontiguousarray(H, dtype=np.float64)
    Z = np.empty_like(H)
    scratch = np.empty_like(H)
    lib.oracle_appnp_f64(n, _ptr(indptr), _ptr(indices), _ptr(val), _ptr(H), _ptr(Z), _ptr(scratch), F, K, alpha)
    return Z
