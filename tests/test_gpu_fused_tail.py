"""The encoder tail fused with the propagation's input scaling (csrc/encoder_tail.cu + PPNP_MODE_SYM_Y0; SURVEY.md
section 8f rank 2): against the PyTorch encoder + the fp64 oracle APPNP, forward and gradients, at 1e-5."""
import os

import numpy as np
import pytest
import torch

from util import load_std, oracle, relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("n,hidden,C,bias,scaled", [(1000, 64, 7, False, True), (4097, 64, 3, True, True), (33, 17, 1, True, False),
                                                      (5000, 256, 64, False, True), (2810, 64, 7, False, False), (64, 1, 5, True, True)])
def test_linear_rowscale_forward_and_adjoint(n, hidden, C, bias, scaled):
    import ppnp_b200 as P
    g = torch.Generator(device=dev()).manual_seed(n)
    A = torch.randn(n, hidden, device=dev(), generator=g)
    W = torch.randn(C, hidden, device=dev(), generator=g)
    b = torch.randn(C, device=dev(), generator=g) if bias else None
    s = (torch.rand(n, device=dev(), generator=g) + 0.5) if scaled else None
    out = P.linear_rowscale(A, W, b, s)
    ref = A.double() @ W.double().t()
    if bias:
        ref = ref + b.double()
    if scaled:
        ref = ref * s.double()[:, None]
    assert relerr(out.cpu().numpy(), ref.cpu().numpy()) < 1e-6
    dOut = torch.randn(n, C, device=dev(), generator=g)
    dA, dW, db = P.linear_rowscale_backward(A, dOut, W, s, need_dA=True, need_dbias=bias)
    gs = dOut.double() * (s.double()[:, None] if scaled else 1.0)
    assert relerr(dA.cpu().numpy(), (gs @ W.double()).cpu().numpy()) < 1e-6
    assert relerr(dW.cpu().numpy(), (gs.t() @ A.double()).cpu().numpy()) < 1e-5
    if bias:
        assert relerr(db.cpu().numpy(), gs.sum(0).cpu().numpy()) < 1e-5
    dA2, dW2, _ = P.linear_rowscale_backward(A, dOut, W, s, need_dA=True, need_dbias=bias)
    assert torch.equal(dW, dW2) and torch.equal(dA, dA2)          # fixed-order reduction


@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
@pytest.mark.parametrize("order", ["natural", "degree"])
def test_scaled_input_propagation_matches_oracle(name, order):
    import ppnp_b200 as P
    z, adj = load_std(name)
    ahat = P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev()), torch.from_numpy(z["adj_indices"]).to(dev()))
    g = P.PropagationGraph(ahat, chunk_edges=128, order=order, keep_vals=False)       # no stored values at all
    H = np.random.RandomState(0).randn(ahat.n, 7).astype(np.float32)
    A = oracle.calc_A_hat(adj, "sym")
    for K in (1, 2, 10):
        Y0 = torch.from_numpy(H).to(dev()) * ahat.dinv[:, None]
        Z = P.appnp_propagate(g, Y0, K, 0.1, scaled_input=True).cpu().numpy()
        assert relerr(Z, oracle.appnp(A, H.astype(np.float64), 0.1, K)) < 1e-5


def test_fused_tail_matches_encoder_plus_oracle_with_gradients():
    import ppnp_b200 as P
    z, adj = load_std("cora_ml")
    ahat = P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev()), torch.from_numpy(z["adj_indices"]).to(dev()))
    g = P.PropagationGraph(ahat, chunk_edges=128, order="degree")
    n, hidden, C, K = ahat.n, 64, 7, 10
    gen = torch.Generator(device=dev()).manual_seed(0)
    A1 = torch.relu(torch.randn(n, hidden, device=dev(), generator=gen)).requires_grad_(True)
    lin = torch.nn.Linear(hidden, C, bias=True).to(dev())
    Zf = P.appnp_fused_tail(A1, lin.weight, lin.bias, g, K, 0.1)
    up = torch.randn(n, C, device=dev(), generator=gen)
    Zf.backward(up)
    gA, gW, gb = A1.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()
    # PyTorch encoder tail + oracle propagation (fp64): Z = P(A1 W^T + b); dH = P(up); dA1 = dH W; dW = dH^T A1; db = sum dH
    A = oracle.calc_A_hat(adj, "sym")
    H = (A1.detach().double() @ lin.weight.detach().double().t() + lin.bias.detach().double()).cpu().numpy()
    Zo = oracle.appnp(A, H, 0.1, K)
    dHo = oracle.appnp(A, up.double().cpu().numpy(), 0.1, K)
    assert relerr(Zf.detach().cpu().numpy(), Zo) < 1e-5
    assert (Zf.detach().cpu().numpy().argmax(1) == Zo.argmax(1)).all()
    assert relerr(gA.cpu().numpy(), dHo @ lin.weight.detach().double().cpu().numpy()) < 1e-5
    assert relerr(gW.cpu().numpy(), dHo.T @ A1.detach().double().cpu().numpy()) < 1e-5
    assert relerr(gb.cpu().numpy(), dHo.sum(0)) < 1e-5
    # and against the unfused differentiable path of this package
    A1b = A1.detach().clone().requires_grad_(True)
    lin.zero_grad()
    P.appnp(torch.nn.functional.linear(A1b, lin.weight, lin.bias), g, K, 0.1).backward(up)
    assert relerr(gA.cpu().numpy(), A1b.grad.cpu().numpy()) < 1e-5 and relerr(gW.cpu().numpy(), lin.weight.grad.cpu().numpy()) < 1e-5


def test_fused_tail_refuses_weighted_graphs():
    import ppnp_b200 as P
    import scipy.sparse as sp
    rng = np.random.RandomState(0)
    d = (rng.rand(200, 200) < 0.05) * rng.rand(200, 200)
    adj = sp.csr_matrix((d + d.T).astype(np.float32)); adj.sort_indices()
    ahat = P.csr_normalize(torch.from_numpy(adj.indptr).to(dev()), torch.from_numpy(adj.indices).to(dev()), torch.from_numpy(adj.data).to(dev()))
    g = P.PropagationGraph(ahat)
    with pytest.raises(ValueError):
        P.appnp_fused_tail(torch.zeros(200, 8, device=dev()), torch.zeros(3, 8, device=dev()), None, g)
    with pytest.raises(ValueError):
        P.appnp_propagate(g, torch.zeros(200, 3, device=dev()), 10, 0.1, scaled_input=True)
