#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- run the REFERENCE (model.PPNP / helpers.compute_ppr imported from
/root/reference, unmodified, on CPU) through the same training recipe and the same frozen split as
tools/train_cora.py, and store the accuracies in tests/golden/cora_ml_train_ref.json.  Build
container only (the GPU box has no reference checkout)."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from helpers import SimpleEarlyStopping, compute_ppr, set_seeds  # noqa: E402 (reference)
from model import PPNP  # noqa: E402 (reference)

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
z = np.load(os.path.join(ROOT, "tests", "golden", "cora_ml_std.npz"))
g = np.load(os.path.join(ROOT, "tests", "golden", "cora_ml_golden.npz"))
n = len(z["adj_indptr"]) - 1
adj = sp.csr_matrix((np.ones(len(z["adj_indices"]), np.float32), z["adj_indices"], z["adj_indptr"]), shape=(n, n))
attr = sp.csr_matrix((z["attr_data"], z["attr_indices"], z["attr_indptr"]), shape=tuple(z["attr_shape"]))
rs = np.asarray(attr.sum(1)).ravel()
X = torch.FloatTensor(np.asarray(attr.multiply(1 / np.maximum(rs, 1e-12)[:, None]).todense()))
y = torch.LongTensor(z["labels"])
idx_train, idx_stop, idx_valid = (torch.LongTensor(g[k]) for k in ("idx_train", "idx_stop", "idx_valid"))
y_train, y_stop, y_valid = y[idx_train], y[idx_stop], y[idx_valid]
set_seeds(123)
ppr = torch.FloatTensor(compute_ppr(adj, alpha=0.1))
records = []
for run in range(runs):
    torch.manual_seed(1000 + run)
    model = PPNP(n_features=X.shape[1], n_classes=y.max() + 1, ppr=ppr)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    es = SimpleEarlyStopping(model)
    t = time.time()
    for epoch in range(10000):
        model.train()
        loss = F.cross_entropy(model(X, idx_train), y_train) + 5e-3 / 2 * model.get_norm()
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
    rec = dict(es.record); rec["run"] = run; rec["ms_per_epoch"] = 1e3 * (time.time() - t) / (epoch + 1)
    records.append(rec)
    print(json.dumps(rec), flush=True)
va = np.array([r["valid_acc"] for r in records])
out = {"what": "reference (CPU) on the frozen Cora-ML split, torch seeds 1000+run", "runs": records,
       "valid_acc_mean": float(va.mean()), "valid_acc_std": float(va.std(ddof=1)) if len(va) > 1 else 0.0}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "cora_ml_train_ref.json"), "w"), indent=1)
print(json.dumps(out))
