"""The drop-in ``model`` / ``helpers`` modules (ppnp_b200/shim).  CPU tests cover the interface
and the training-harness utilities (against the reference itself when /root/reference exists);
GPU tests run the reference's call pattern (main.py:106-121, batch-main.py:111-146) end to end."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from util import ROOT, load_golden, load_std, oracle, relerr

SHIM = os.path.join(ROOT, "ppnp_b200", "shim")
REF = "/root/reference"


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def shim(monkeypatch):
    """Import the shim modules by bare name, the way main.py does."""
    monkeypatch.syspath_prepend(SHIM)
    for m in ("helpers", "model"):
        sys.modules.pop(m, None)
    import helpers
    import model
    yield helpers, model
    for m in ("helpers", "model"):
        sys.modules.pop(m, None)


def test_exports_and_signatures(shim):
    helpers, model = shim
    for name in ("set_seeds", "compute_ppr", "SimpleEarlyStopping", "calc_A_hat"):
        assert hasattr(helpers, name)
    import inspect
    assert list(inspect.signature(helpers.compute_ppr).parameters) == ["adj", "alpha", "mode"]
    assert list(inspect.signature(model.PPNP.__init__).parameters) == \
        ["self", "n_features", "n_classes", "ppr", "hidden_dim", "drop_prob", "bias"]
    assert list(inspect.signature(model.PPNP.forward).parameters) == ["self", "X", "idx", "ppr"]


def test_ppnp_module_surface_cpu(shim):
    _, model = shim
    ppr = torch.eye(5)
    m = model.PPNP(n_features=4, n_classes=torch.tensor(2) + 1, ppr=ppr)      # 0-d LongTensor like main.py:107
    assert [type(l).__name__ for l in m.encoder] == ["Dropout", "CustomLinear", "ReLU", "Dropout", "Linear"]
    assert set(m.state_dict()) == {"encoder.1.weight", "encoder.4.weight", "ppr"}
    assert m.encoder[1].weight.shape == (4, 64) and m.encoder[4].weight.shape == (3, 64)
    assert float(m.get_norm()) == pytest.approx(float((m.encoder[1].weight ** 2).sum()))
    with pytest.raises(Exception):
        m(torch.zeros(5, 4))                                                 # model.py:67
    # on the CPU the buffer indexes like a plain tensor (batch-main.py:140-142 semantics)
    sub = m.ppr[torch.tensor([0, 2])]
    assert torch.equal(torch.as_tensor(sub), ppr[[0, 2]])
    # and the CUDA path refuses to fall back
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(5, 4), torch.tensor([0, 1]))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_encoder_init_stream_matches_reference(shim):
    _, model = shim
    ref_model = load_module("ref_model", os.path.join(REF, "model.py"))
    torch.manual_seed(4217546909)
    a = ref_model.PPNP(n_features=30, n_classes=7, ppr=torch.eye(3))
    torch.manual_seed(4217546909)
    b = model.PPNP(n_features=30, n_classes=7, ppr=torch.eye(3))
    for k in ("encoder.1.weight", "encoder.4.weight"):
        assert torch.equal(a.state_dict()[k], b.state_dict()[k])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_early_stopping_and_seeds_match_reference(shim):
    helpers, _ = shim
    ref = load_module("ref_helpers", os.path.join(REF, "helpers.py"))
    rng = np.random.RandomState(0)
    accs = np.round(rng.rand(400) * 0.2 + np.linspace(0.5, 0.8, 400), 3)
    losses = np.round(rng.rand(400) * 0.3 + np.linspace(1.5, 0.6, 400), 3)
    accs[250:] = 0.3
    losses[250:] = 3.0
    a = ref.SimpleEarlyStopping(None, patience=20)
    b = helpers.SimpleEarlyStopping(None, patience=20)
    for e, (acc, loss) in enumerate(zip(accs, losses)):
        ra = a.should_stop(float(acc), float(loss), e, record={"epoch": e})
        rb = b.should_stop(float(acc), float(loss), e, record={"epoch": e})
        assert ra == rb and a.patience == b.patience and a.best_epoch == b.best_epoch and a.record == b.record
        if ra:
            break
    assert ra
    import random
    ref.set_seeds(123); x = (random.random(), np.random.rand(), torch.rand(1).item())
    helpers.set_seeds(123); y = (random.random(), np.random.rand(), torch.rand(1).item())
    assert x == y


# --------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
def test_calc_A_hat_drop_in(shim, name):
    helpers, _ = shim
    _, adj = load_std(name)
    g = load_golden(name)
    for mode in ("sym", "rw"):
        ah = helpers.calc_A_hat(adj, mode)
        assert ah.dtype == np.float64 and ah.format == "csr"
        assert np.array_equal(ah.indptr, g[f"ahat_{mode}_indptr"])
        assert np.array_equal(ah.indices, g[f"ahat_{mode}_indices"])
        assert np.array_equal(ah.data, g[f"ahat_{mode}_data"])


@pytest.mark.gpu
@pytest.mark.parametrize("gemm", ["fp32", "bf16"])
def test_main_py_call_pattern_exact(shim, monkeypatch, gemm):
    """main.py:106-107,121,138: FloatTensor(compute_ppr(...)) -> PPNP(...).cuda() -> model(X, idx)."""
    helpers, model = shim
    monkeypatch.setenv("PPNP_MODE", "exact")
    monkeypatch.setenv("PPNP_GEMM", gemm)
    z, adj = load_std("cora_ml")
    g = load_golden("cora_ml")
    out = helpers.compute_ppr(adj, alpha=0.1)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (adj.shape[0],) * 2
    ppr = torch.FloatTensor(out)                                   # main.py:106
    assert relerr(ppr[torch.from_numpy(g["ppr_rows_idx"])].numpy(), g["ppr_rows"]) < 1e-5
    torch.manual_seed(1234)
    m = model.PPNP(n_features=2879, n_classes=torch.tensor(6) + 1, ppr=ppr).cuda()
    m.eval()
    import scipy.sparse as sp
    attr = sp.csr_matrix((z["attr_data"], z["attr_indices"], z["attr_indptr"]), shape=tuple(z["attr_shape"]))
    rs = np.asarray(attr.sum(1)).ravel()
    X = torch.FloatTensor(np.asarray(attr.multiply(1 / np.maximum(rs, 1e-12)[:, None]).todense())).cuda()
    idx = torch.from_numpy(g["idx_train"]).cuda()
    with torch.no_grad():
        H = m.encoder(X)
        logits = m(X, idx)
    ref = oracle.ppnp_forward(oracle.compute_ppr(adj, 0.1).astype(np.float32), H.cpu().numpy(), g["idx_train"])
    tol = 1e-5 if gemm == "fp32" else 1e-2
    assert relerr(logits.cpu().numpy(), ref) < tol
    if gemm == "fp32":
        assert (logits.cpu().numpy().argmax(1) == ref.argmax(1)).all()
    # one training step runs (autograd through the CUDA op) and changes the weights
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    y = torch.from_numpy(z["labels"][g["idx_train"]]).cuda()
    w0 = m.encoder[1].weight.detach().clone()
    loss = torch.nn.functional.cross_entropy(m(X, idx), y) + 5e-3 / 2 * m.get_norm()
    opt.zero_grad(); loss.backward(); opt.step()
    assert torch.isfinite(loss) and not torch.equal(w0, m.encoder[1].weight)


@pytest.mark.gpu
def test_main_py_call_pattern_appnp(shim, monkeypatch):
    helpers, model = shim
    monkeypatch.setenv("PPNP_MODE", "appnp")
    monkeypatch.setenv("PPNP_K", "10")
    _, adj = load_std("citeseer")
    g = load_golden("citeseer")
    ppr = torch.FloatTensor(helpers.compute_ppr(adj, alpha=0.1))
    torch.manual_seed(7)
    m = model.PPNP(n_features=16, n_classes=6, ppr=ppr).cuda()
    m.eval()
    X = torch.randn(adj.shape[0], 16).cuda()
    idx = torch.from_numpy(g["idx_stop"]).cuda()
    with torch.no_grad():
        H = m.encoder(X)
        out = m(X, idx)
    ref = oracle.appnp(oracle.calc_A_hat(adj, "sym"), H.cpu().numpy().astype(np.float64), 0.1, 10)[g["idx_stop"]]
    assert relerr(out.cpu().numpy(), ref) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("batch_mode", ["sparse", "dense"])
def test_batch_main_call_pattern(shim, monkeypatch, batch_mode):
    """batch-main.py:111-119 and 140-152 written out literally against the shim."""
    helpers, model = shim
    monkeypatch.setenv("PPNP_MODE", "exact")
    monkeypatch.setenv("PPNP_GEMM", "fp32")
    monkeypatch.setenv("PPNP_BATCH", batch_mode)
    _, adj = load_std("citeseer")
    n = adj.shape[0]
    ppr = torch.FloatTensor(helpers.compute_ppr(adj, alpha=0.1))
    thresh, _ = ppr.topk(32, axis=-1)                       # batch-main.py:115
    ppr[ppr < thresh[:, -1]] = 0                             # :116
    torch.manual_seed(3)
    m = model.PPNP(n_features=20, n_classes=6, ppr=ppr).cuda()
    X = torch.randn(n, 20).cuda()
    idx_batch = torch.from_numpy(np.sort(np.random.RandomState(0).choice(n, 64, replace=False)))   # CPU index, as the DataLoader yields
    y_batch = torch.randint(0, 6, (64,)).cuda()
    m.eval()
    ppr_sub = m.ppr[idx_batch]                               # :140
    sel = (ppr_sub > 0).any(dim=0)                           # :141
    ppr_sub = ppr_sub[:, sel]                                # :142
    X_batch = X[sel].cuda()                                  # :144
    logits = m(X_batch, idx=None, ppr=ppr_sub)               # :146
    loss = torch.nn.functional.cross_entropy(logits, y_batch) + 5e-3 / 2 * m.get_norm()
    loss.backward()
    with torch.no_grad():
        Hfull = m.encoder(X).cpu().numpy().astype(np.float64)
    ref_logits, ref_sel = oracle.batch_step(ppr.numpy(), idx_batch.numpy(), Hfull)
    assert np.array_equal(sel.cpu().numpy(), ref_sel)
    assert relerr(logits.detach().cpu().numpy(), ref_logits) < 1e-5
    assert m.encoder[4].weight.grad is not None and torch.isfinite(m.encoder[4].weight.grad).all()
    # eval path of batch-main.py:161-169 goes through model.py:63 on the sparsified buffer
    with torch.no_grad():
        ev = m(X, torch.arange(100).cuda())
    assert relerr(ev.cpu().numpy(), ppr.numpy()[:100].astype(np.float64) @ Hfull) < 1e-5


# ------------------------------------------------------------------ ppnp.data.sparsegraph overlay
REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs the reference checkout (build container only)")
def test_sparsegraph_overlay_resolves_like_main_py(monkeypatch):
    """main.py:27-28 imports with [shim, reference] on sys.path: SparseGraph is the GPU-backed subclass,
    ppnp.preprocessing and the data files still come from the reference, and there is no CPU fallback."""
    saved = {k: sys.modules.pop(k) for k in [k for k in sys.modules if k == "ppnp" or k.startswith("ppnp.")]}
    monkeypatch.setattr(sys, "path", [SHIM, REFERENCE] + [p for p in sys.path if p not in (SHIM, REFERENCE)])
    try:
        _overlay_checks()
    finally:                              # leave no overlay modules behind for the tests that follow
        for k in [k for k in sys.modules if k == "ppnp" or k.startswith("ppnp.") or k == "_ppnp_reference_sparsegraph"]:
            sys.modules.pop(k, None)
        sys.modules.update(saved)


def _overlay_checks():
    from ppnp.data.sparsegraph import SparseGraph, create_subgraph, largest_connected_components  # noqa: F401
    from ppnp.preprocessing import gen_splits, normalize_attributes  # noqa: F401
    import ppnp.preprocessing as prep
    import ppnp.data.sparsegraph as sgm
    assert prep.__file__.startswith(REFERENCE) and sgm.__file__.startswith(SHIM)
    raw = np.load(os.path.join(REFERENCE, "ppnp", "data", "citeseer.npz"), allow_pickle=True)
    g = SparseGraph.from_flat_dict(dict(raw))
    assert type(g) is SparseGraph and type(g).__mro__[1].__module__ == "_ppnp_reference_sparsegraph"
    assert g.num_nodes() == 3312
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            g.standardize(select_lcc=True)
    else:
        g.standardize(select_lcc=True)
        assert g.num_nodes() == 2110 and g.adj_matrix.nnz == 7336
