#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz by importing and running the REFERENCE
(/root/reference, read-only) unmodified.  Run once in the build container:

    python oracle/gen_golden.py

The GPU box has no /root/reference; tests there read only the committed .npz files.  Nothing in
this script is used at test time.  What is frozen (per data set shipped with the reference):

  <name>_std.npz     the hot path's INPUT: SparseGraph.from_flat_dict + standardize(select_lcc=True)
                     (main.py:73-75): adjacency CSR (int32), attribute CSR, labels
  <name>_golden.npz  outputs of the reference functions on that input:
      helpers.calc_A_hat 'sym' and 'rw'        (helpers.py:58-66)  indptr, indices, data fp64
      helpers.compute_ppr alpha=0.1            (helpers.py:68-71)  16 full rows, diagonal, row sums
      model.PPNP.forward(X, idx)               (model.py:61-63)    encoder output H, logits, autograd dH
      batch-main.py:115-116 top-k sparsify     k = 32              thresholds, per-row/col nnz, 16 rows
      batch-main.py:140-146 one batch          B = 32              sel, logits
      APPNP K=10 restated on the reference's A_hat (fp64)          Z_10, and its K=300 limit vs ppr @ H
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from helpers import calc_A_hat, compute_ppr  # noqa: E402  (reference)
from model import PPNP  # noqa: E402  (reference)
from ppnp.data.sparsegraph import SparseGraph  # noqa: E402  (reference)
from ppnp.preprocessing import gen_splits, normalize_attributes  # noqa: E402  (reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
ALPHA = 0.1
K_TOP = 32
BATCH = 32


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def one(name):
    raw = np.load(os.path.join(REF, "ppnp", "data", f"{name}.npz"), allow_pickle=True)
    graph = SparseGraph.from_flat_dict(dict(raw))
    graph.standardize(select_lcc=True)
    adj = graph.adj_matrix.tocsr()
    adj.sort_indices()
    attr = graph.attr_matrix.tocsr()
    labels = graph.labels
    n = adj.shape[0]
    np.savez_compressed(os.path.join(OUT, f"{name}_std.npz"),
                        adj_indptr=adj.indptr.astype(np.int32), adj_indices=adj.indices.astype(np.int32),
                        attr_indptr=attr.indptr.astype(np.int32), attr_indices=attr.indices.astype(np.int32),
                        attr_data=attr.data.astype(np.float32), attr_shape=np.array(attr.shape), labels=labels.astype(np.int64))

    g = {}
    # ---- helpers.calc_A_hat
    for mode in ("sym", "rw"):
        ah = calc_A_hat(adj, mode).tocsr()
        ah.sort_indices()
        g[f"ahat_{mode}_indptr"] = ah.indptr.astype(np.int32)
        g[f"ahat_{mode}_indices"] = ah.indices.astype(np.int32)
        g[f"ahat_{mode}_data"] = ah.data.astype(np.float64)
    ah = calc_A_hat(adj, "sym").tocsr()
    print(name, "n", n, "nnz(A)", adj.nnz, "nnz(A_hat)", ah.nnz, "sha16 indptr/indices/data",
          sha16(ah.indptr.astype(np.int32)), sha16(ah.indices.astype(np.int32)), sha16(ah.data))

    # ---- helpers.compute_ppr
    ppr64 = compute_ppr(adj, ALPHA)
    rng = np.random.RandomState(7)
    rows = np.sort(rng.choice(n, 16, replace=False))
    g["ppr_rows_idx"] = rows.astype(np.int64)
    g["ppr_rows"] = ppr64[rows]
    g["ppr_diag"] = np.diag(ppr64).copy()
    g["ppr_rowsum"] = ppr64.sum(1)
    g["ppr_fro"] = np.array(np.linalg.norm(ppr64))

    # ---- model.PPNP.forward on the literal main.py objects
    X = normalize_attributes(graph.attr_matrix)
    X = torch.FloatTensor(np.asarray(X.todense()))
    y = torch.LongTensor(labels)
    idx_split_args = {"ntrain_per_class": 20, "nstopping": 500, "nknown": 1500, "seed": 2413340114}
    idx_train, idx_stop, idx_valid = gen_splits(labels, idx_split_args, test=False)
    g["idx_train"], g["idx_stop"], g["idx_valid"] = (np.asarray(a, dtype=np.int64) for a in (idx_train, idx_stop, idx_valid))
    ppr = torch.FloatTensor(ppr64)
    torch.manual_seed(1234)
    model = PPNP(n_features=X.shape[1], n_classes=y.max() + 1, ppr=ppr)
    model.eval()
    H = model.encoder(X).detach().clone().requires_grad_(True)
    idx = torch.LongTensor(idx_train)
    logits = model.ppr[idx] @ H                      # model.py:63 with the encoder output frozen
    with torch.no_grad():
        assert torch.equal(logits, model(X, idx)), "frozen-H forward must equal PPNP.forward"
    G = torch.from_numpy(np.random.RandomState(11).randn(*logits.shape).astype(np.float32))
    logits.backward(G)
    g["H"] = H.detach().numpy()
    g["logits_train"] = logits.detach().numpy()
    g["G_train"] = G.numpy()
    g["dH_train"] = H.grad.numpy()
    with torch.no_grad():
        g["logits_full"] = (model.ppr @ H).numpy()   # model.py:65 with ppr = the buffer

    # ---- batch-main.py:115-116 and 140-146 (literal lines on CPU tensors)
    ppr_b = torch.FloatTensor(ppr64)
    thresh, _ = ppr_b.topk(K_TOP, axis=-1)
    ppr_b[ppr_b < thresh[:, -1]] = 0
    g["topk_k"] = np.array(K_TOP)
    g["topk_thresh"] = thresh[:, -1].numpy().copy()
    g["topk_row_nnz"] = (ppr_b > 0).sum(1).numpy().astype(np.int32)
    g["topk_col_nnz"] = (ppr_b > 0).sum(0).numpy().astype(np.int32)
    g["topk_rows"] = ppr_b[torch.from_numpy(rows)].numpy()
    model_b = PPNP(n_features=X.shape[1], n_classes=y.max() + 1, ppr=ppr_b)
    model_b.load_state_dict({**model.state_dict(), "ppr": ppr_b})
    model_b.eval()
    idx_batch = torch.LongTensor(np.sort(np.random.RandomState(3).choice(n, BATCH, replace=False)))
    with torch.no_grad():
        ppr_sub = model_b.ppr[idx_batch]             # batch-main.py:140
        sel = (ppr_sub > 0).any(dim=0)               # :141
        ppr_sub = ppr_sub[:, sel]                    # :142
        X_batch = X[sel]                             # :144
        logits_b = model_b(X_batch, idx=None, ppr=ppr_sub)  # :146
    g["batch_idx"] = idx_batch.numpy()
    g["batch_sel"] = sel.numpy()
    g["batch_logits"] = logits_b.numpy()

    # ---- APPNP K=10 restated on the reference's A_hat (fp64), and the K -> inf limit (KAT-1)
    Hn = g["H"].astype(np.float64)
    Z = Hn.copy()
    for _ in range(10):
        Z = (1 - ALPHA) * (ah @ Z) + ALPHA * Hn
    g["appnp_K10"] = Z
    Zl = Hn.copy()
    for _ in range(300):
        Zl = (1 - ALPHA) * (ah @ Zl) + ALPHA * Hn
    lim = ppr64 @ Hn
    g["appnp_limit_relerr"] = np.array(np.linalg.norm(Zl - lim) / np.linalg.norm(lim))
    print(name, "KAT-1 |appnp_K300 - ppr@H| / |ppr@H| =", float(g["appnp_limit_relerr"]))
    g["alpha"] = np.array(ALPHA)
    np.savez_compressed(os.path.join(OUT, f"{name}_golden.npz"), **g)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for nm in ("cora_ml", "citeseer"):
        one(nm)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
