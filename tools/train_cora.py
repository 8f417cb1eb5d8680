#!/usr/bin/env python
"""Accuracy check of the drop-in modules on the reference's own graph: the training recipe of
main.py (Adam lr 0.01, CE + 5e-3/2 |W1|^2, early stopping on the stopping set, patience 100) on the
frozen Cora-ML fixture (tests/golden/cora_ml_std.npz, default split of ppnp/preprocessing.py) with
`model` / `helpers` imported by bare name from ppnp_b200/shim, in exact and APPNP mode.  Reference
numbers for the same split come from the reference run on CPU (tests/golden/cora_ml_train_ref.json).
    python tools/train_cora.py [--mode exact|appnp] [--runs 5] [--gemm fp32|bf16]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="exact")
    ap.add_argument("--gemm", default="fp32")
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--max-epochs", type=int, default=10000)
    ap.add_argument("--dataset", default="cora_ml")
    args = ap.parse_args()
    os.environ["PPNP_MODE"], os.environ["PPNP_GEMM"] = args.mode, args.gemm
    sys.path[:0] = [os.path.join(ROOT, "ppnp_b200", "shim"), ROOT]
    from helpers import SimpleEarlyStopping, compute_ppr, set_seeds   # the drop-in modules
    from model import PPNP

    z = np.load(os.path.join(ROOT, "tests", "golden", f"{args.dataset}_std.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{args.dataset}_golden.npz"))
    n = len(z["adj_indptr"]) - 1
    adj = sp.csr_matrix((np.ones(len(z["adj_indices"]), np.float32), z["adj_indices"], z["adj_indptr"]), shape=(n, n))
    attr = sp.csr_matrix((z["attr_data"], z["attr_indices"], z["attr_indptr"]), shape=tuple(z["attr_shape"]))
    rs = np.asarray(attr.sum(1)).ravel()
    X = torch.FloatTensor(np.asarray(attr.multiply(1 / np.maximum(rs, 1e-12)[:, None]).todense())).cuda()
    y = torch.LongTensor(z["labels"])
    idx_train, idx_stop, idx_valid = (torch.LongTensor(g[k]).cuda() for k in ("idx_train", "idx_stop", "idx_valid"))
    y_train, y_stop, y_valid = y[idx_train.cpu()].cuda(), y[idx_stop.cpu()].cuda(), y[idx_valid.cpu()].cuda()
    set_seeds(123)
    records = []
    for run in range(args.runs):
        torch.manual_seed(1000 + run)
        ppr = torch.FloatTensor(compute_ppr(adj, alpha=0.1))
        model = PPNP(n_features=X.shape[1], n_classes=y.max() + 1, ppr=ppr).cuda()
        opt = torch.optim.Adam(model.parameters(), lr=0.01)
        es = SimpleEarlyStopping(model)
        t = time.time()
        for epoch in range(args.max_epochs):
            model.train()
            logits = model(X, idx_train)
            loss = F.cross_entropy(logits, y_train) + 5e-3 / 2 * model.get_norm()
            opt.zero_grad(); loss.backward(); opt.step()
            model.eval()
            with torch.no_grad():
                ls = model(X, idx_stop)
                stop_loss = F.cross_entropy(ls, y_stop) + 5e-3 / 2 * model.get_norm()
                stop_acc = (ls.argmax(-1) == y_stop).float().mean()
                valid_acc = (model(X, idx_valid).argmax(-1) == y_valid).float().mean()
            rec = {"epoch": epoch, "elapsed": time.time() - t, "stop_acc": float(stop_acc), "valid_acc": float(valid_acc)}
            if es.should_stop(acc=float(stop_acc), loss=float(stop_loss), epoch=epoch, record=rec):
                break
        rec = dict(es.record); rec["run"] = run; rec["epochs_run"] = epoch + 1; rec["ms_per_epoch"] = 1e3 * (time.time() - t) / (epoch + 1)
        records.append(rec)
        print(json.dumps(rec), flush=True)
    va = np.array([r["valid_acc"] for r in records]); sa = np.array([r["stop_acc"] for r in records])
    print(json.dumps({"summary": True, "mode": args.mode, "gemm": args.gemm, "dataset": args.dataset, "runs": args.runs,
                      "valid_acc_mean": float(va.mean()), "valid_acc_std": float(va.std(ddof=1)) if len(va) > 1 else 0.0,
                      "stop_acc_mean": float(sa.mean()), "ms_per_epoch": float(np.mean([r["ms_per_epoch"] for r in records]))}))


if __name__ == "__main__":
    main()
