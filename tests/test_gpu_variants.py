"""Plan-selected variants of the propagation kernel (include/ppnp_b200.h PPNP_PLAN_*): carved streams
on 1024-thread CTAs and 16-byte staging of lane-transposed streams, against the same oracle and
golden fixtures as the default path.  Needs a GPU."""
import numpy as np
import pytest
import torch

from util import load_golden, load_std, oracle, relerr

pytestmark = pytest.mark.gpu

VARIANTS = {
    "degree+idx16": dict(order="degree", idx16=True),
    "carve": dict(order="carve", carve=dict(block_cols=128, n_blocks=16, min_piece=3)),
    "carve+idx16": dict(order="carve", idx16=True, carve=dict(block_cols=128, n_blocks=16, min_piece=3)),
    "carve-narrow+idx16": dict(order="carve", idx16=True, carve=dict(block_cols=64, n_blocks=8, min_piece=2, wide_cta=False)),
    "window+idx16": dict(order="window", idx16=True),
    "window-wide+idx16": dict(order="window", idx16=True, window=dict(key="first", wide_cta=True)),
}


def dev():
    return torch.device("cuda:0")


def gpu_ahat(name):
    import ppnp_b200 as P
    z, adj = load_std(name)
    return P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev()), torch.from_numpy(z["adj_indices"]).to(dev())), adj


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variants_on_skewed_rmat_vs_c_oracle(variant):
    """Hub rows, isolated rows, F = 64 (16-lane groups) and F = 16 (4-lane groups), both value forms."""
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(50000, 1200000, 16, seed=0)
    oip, oidx, oval, odeg = oracle.c_a_hat(ip, idx, None, "sym")
    ahat = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()))
    graph = P.PropagationGraph(ahat, chunk_edges=256, **VARIANTS[variant])
    assert graph.plan.n_chunks > 4096          # the per-step kernels, not the one-launch K-step kernel
    if "carve" in variant:
        assert graph.plan.carve["carved_edges"] > 0 and graph.plan.n_fix > 1000
    for F in (64, 16):
        if graph.idx16:
            assert graph.plan_for(F).lane_group == (16 if F == 64 else 4)
        Hn = np.random.RandomState(F).randn(50000, F).astype(np.float32)
        ref = oracle.c_appnp_f64(oip, oidx, oval, Hn.astype(np.float64), 10, 0.1)
        for use_vals in (False, True):
            Z = P.appnp_propagate(graph, torch.from_numpy(Hn).to(dev()), K=10, alpha=0.1, use_vals=use_vals).cpu().numpy()
            assert relerr(Z, ref) < 1e-5, (variant, F, use_vals)
    # a width without a lane-transposed kernel keeps the linear stream (and the same answer)
    Hn = np.random.RandomState(7).randn(50000, 32).astype(np.float32)
    ref = oracle.c_appnp_f64(oip, oidx, oval, Hn.astype(np.float64), 3, 0.1)
    Z = P.appnp_propagate(graph, torch.from_numpy(Hn).to(dev()), K=3, alpha=0.1).cpu().numpy()
    assert relerr(Z, ref) < 1e-5


@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variants_match_golden_appnp(name, variant):
    import ppnp_b200 as P
    ahat, adj = gpu_ahat(name)
    g = load_golden(name)
    kw = dict(VARIANTS[variant])
    if "carve" in kw:
        kw["carve"] = dict(kw["carve"], block_cols=32, n_blocks=8)
    graph = P.PropagationGraph(ahat, chunk_edges=128, **kw)
    A = oracle.calc_A_hat(adj, "sym")
    # one step through ppnp_spmm_step (never the one-launch kernel), F = 64 and 16
    for F in (64, 16):
        Hn = np.random.RandomState(F).randn(adj.shape[0], F).astype(np.float32)
        H = torch.from_numpy(Hn).to(dev())
        Z = P.spmm_step(graph, H, H, 0.1, epi=0, use_vals=True).cpu().numpy()
        assert relerr(Z, oracle.appnp(A, Hn.astype(np.float64), 0.1, 1)) < 1e-6, (variant, F)
    # the golden K = 10 logits (F = number of classes: linear stream, default kernels, carved order)
    H = torch.from_numpy(g["H"]).to(dev())
    Z = P.appnp_propagate(graph, H, K=10, alpha=0.1).cpu().numpy()
    assert relerr(Z, g["appnp_K10"]) < 1e-5
    assert (Z.argmax(1) == g["appnp_K10"].argmax(1)).all()


def test_variants_are_deterministic_and_agree_with_the_default_order():
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(50000, 1200000, 16, seed=1)
    ahat = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()))
    H = torch.randn(50000, 64, device=dev())
    gb = P.PropagationGraph(ahat, order="degree")
    assert gb.plan.n_chunks > 4096          # per-step kernels on both sides (the one-launch kernel adds in another order)
    base = P.appnp_propagate(gb, H, 10, 0.1)
    g16 = P.PropagationGraph(ahat, order="degree", idx16=True)
    a = P.appnp_propagate(g16, H, 10, 0.1)
    assert torch.equal(a, base)                       # same additions in the same order, only the staging differs
    gc = P.PropagationGraph(ahat, order="carve", idx16=True, carve=dict(block_cols=128, n_blocks=16, min_piece=3))
    c1 = P.appnp_propagate(gc, H, 10, 0.1)
    c2 = P.appnp_propagate(gc, H, 10, 0.1)
    assert torch.equal(c1, c2)
    assert float((c1 - base).norm() / base.norm()) < 1e-5


def test_lane_transposed_plan_rejects_other_widths():
    import ppnp_b200 as P
    from ppnp_b200 import _lib
    from ppnp_b200.plan import lane_transpose
    ahat, adj = gpu_ahat("citeseer")
    graph = P.PropagationGraph(ahat, chunk_edges=128, order="degree")
    p16 = lane_transpose(graph.plan, 16)
    lib = _lib.load()
    for F in (16, 48, 128):      # G = 4, a partial tile of G = 16, G = 32
        H = torch.randn(adj.shape[0], F, device=dev())
        out = torch.empty_like(H)
        partial = graph.partial_buffer(F)
        rc = lib.ppnp_spmm_step(p16.struct(), _lib.ptr(H), _lib.ptr(H), _lib.ptr(out), _lib.ptr(partial), F, F, 0.1, 0, 1,
                                _lib.current_stream())
        assert rc in (-1, -3) and b"lane" in lib.ppnp_last_error()
    with pytest.raises(RuntimeError):
        H = torch.randn(adj.shape[0], 64, device=dev())
        Z, S = torch.empty_like(H), torch.empty_like(H)
        rc = lib.ppnp_appnp_propagate_persistent(p16.struct(), _lib.ptr(H), _lib.ptr(Z), _lib.ptr(S), _lib.ptr(graph.partial_buffer(64)),
                                                 64, 64, 2, 0.1, _lib.current_stream())
        _lib.check(rc, "ppnp_appnp_propagate_persistent")
