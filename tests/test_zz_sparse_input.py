"""Sparse-input first layer (ppnp_b200.SparseInput / sparse_first_layer): the
encoder's Dropout + CustomLinear pair (model.py:47-48) over X's stored entries, run by the propagation
kernel.  CPU: the two stream plans and the entry permutations, walked in numpy.  GPU (marked; written after
the round's GPU budget was spent, the file sorts last): forward and weight gradient against dense torch."""
import numpy as np
import pytest
import torch

from util import load_std
from test_dist_cpu import walker_step
import ppnp_b200 as P


def attr_csr(name="citeseer", rows=400):
    z, _ = load_std(name)
    ip, idx, val = z["attr_indptr"][: rows + 1].astype(np.int64), z["attr_indices"], z["attr_data"]
    nnz = int(ip[-1])
    ip = ip.copy()
    ip[5:] -= ip[5] - ip[4]          # make row 4 empty (a node without attributes)
    keep = np.r_[0:int(z["attr_indptr"][4]), int(z["attr_indptr"][5]):nnz]
    return ip, idx[keep].astype(np.int32), val[keep].astype(np.float32), int(z["attr_shape"][1])


def test_sparse_input_streams_walk_to_the_dense_products():
    ip, idx, val, F_in = attr_csr()
    n = len(ip) - 1
    sx = P.SparseInput(torch.from_numpy(ip), torch.from_numpy(idx), torch.from_numpy(val), F_in, chunk_edges=128)
    X = np.zeros((n, F_in))
    rows = np.repeat(np.arange(n), np.diff(ip))
    X[rows, idx] = val
    rng = np.random.RandomState(0)
    W, G = rng.randn(F_in, 6).astype(np.float32), rng.randn(n, 6).astype(np.float32)
    scale = torch.from_numpy((rng.rand(len(val)) < 0.5).astype(np.float32) * 2)
    Xd = np.zeros_like(X)
    Xd[rows, idx] = val * scale.numpy()
    for sc, Xm in ((None, X), (scale, Xd)):
        out = torch.zeros(n, 6)
        walker_step(sx._with_values(sx.fwd, sx.fwd_edge, sc), torch.from_numpy(W), out, out, 0.0, 16, True)   # PLAIN | ACC
        assert np.allclose(out.numpy(), Xm @ W, rtol=1e-5, atol=1e-5)
        assert (out[4] == 0).all()                                   # the empty row is not in the stream
        dW = torch.zeros(F_in, 6)
        walker_step(sx._with_values(sx.bwd, sx.bwd_edge, sc), torch.from_numpy(G), dW, dW, 0.0, 16, True)
        assert np.allclose(dW.numpy(), Xm.T @ G, rtol=1e-5, atol=1e-5)
    assert sx.fwd.n == n and sx.bwd.n == F_in and sx.nnz == len(val)
    with pytest.raises(RuntimeError):        # no CPU path for the launch itself
        sx.matmul(torch.from_numpy(W))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
def test_sparse_first_layer_matches_dense_torch(name):
    z, _ = load_std(name)
    dev = torch.device("cuda:0")
    n, F_in = int(z["attr_shape"][0]), int(z["attr_shape"][1])
    Xs = torch.sparse_csr_tensor(torch.from_numpy(z["attr_indptr"].astype(np.int64)), torch.from_numpy(z["attr_indices"].astype(np.int64)),
                                 torch.from_numpy(z["attr_data"]), size=(n, F_in)).to(dev)
    X = Xs.to_dense()
    sx = P.SparseInput.from_dense(X)
    assert sx.nnz == len(z["attr_data"])
    W = torch.randn(F_in, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(0), requires_grad=True)
    G = torch.randn(n, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    # eval mode: no dropout, must equal X @ W (model.py:34-38 addmm) to fp32 round-off
    out = P.sparse_first_layer(sx, W, 0.5, training=False)
    ref = X.double() @ W.detach().double()
    assert float((out.double() - ref).norm() / ref.norm()) < 1e-6
    out.backward(G)
    gref = X.double().T @ G.double()
    assert float((W.grad.double() - gref).norm() / gref.norm()) < 1e-6
    # training mode: one mask for forward and backward, inverted-dropout scaling, about half the entries kept
    W.grad = None
    gen = torch.Generator(device=dev).manual_seed(7)
    out = P.sparse_first_layer(sx, W, 0.5, training=True, generator=gen)
    keep = torch.empty(sx.nnz, device=dev).bernoulli_(0.5, generator=torch.Generator(device=dev).manual_seed(7))
    s = X.to_sparse_csr()
    Xd = torch.sparse_csr_tensor(s.crow_indices(), s.col_indices(), s.values() * keep * 2, size=X.shape).to_dense()
    ref = Xd.double() @ W.detach().double()
    assert float((out.double() - ref).norm() / ref.norm()) < 1e-6
    out.backward(G)
    gref = Xd.double().T @ G.double()
    assert float((W.grad.double() - gref).norm() / gref.norm()) < 1e-6
    assert 0.45 < float(keep.mean()) < 0.55


@pytest.mark.gpu
def test_shim_sparse_x_switch_matches_the_dense_encoder(monkeypatch):
    """PPNP_SPARSE_X=1: model.PPNP.forward(X, idx) (model.py:61-63) with the first layer over X's stored
    entries equals the dense encoder in eval mode, and trains (the weight gradient flows) in train mode."""
    import os
    import sys
    from util import ROOT
    monkeypatch.setenv("PPNP_SPARSE_X", "1")
    monkeypatch.syspath_prepend(os.path.join(ROOT, "ppnp_b200", "shim"))
    for m in ("helpers", "model"):
        sys.modules.pop(m, None)
    import model
    z, _ = load_std("citeseer")
    dev = torch.device("cuda:0")
    n, F_in = int(z["attr_shape"][0]), int(z["attr_shape"][1])
    X = torch.sparse_csr_tensor(torch.from_numpy(z["attr_indptr"].astype(np.int64)), torch.from_numpy(z["attr_indices"].astype(np.int64)),
                                torch.from_numpy(z["attr_data"]), size=(n, F_in)).to_dense().to(dev)
    torch.manual_seed(0)
    ppr = torch.rand(n, n) * (torch.rand(n, n) < 0.01)
    net = model.PPNP(F_in, torch.tensor(6), ppr).to(dev)
    idx = torch.arange(0, n, 7, device=dev)
    net.eval()
    with torch.no_grad():
        a = net(X, idx)
        net._sparse_x = False
        b = net(X, idx)
        net._sparse_x = True
    assert float((a - b).norm() / b.norm()) < 1e-5
    net.train()
    out = net(X, idx)
    out.sum().backward()
    g = net.encoder[1].weight.grad
    assert g is not None and float(g.abs().sum()) > 0 and torch.isfinite(g).all()
    for m in ("helpers", "model"):
        sys.modules.pop(m, None)
