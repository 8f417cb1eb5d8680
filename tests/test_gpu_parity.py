"""Parity of the CUDA path (through the C ABI) with the oracle and the golden fixtures.
Needs a GPU: run with  pytest -m gpu  on the B200 box."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from util import load_golden, load_std, oracle, relerr

pytestmark = pytest.mark.gpu
NAMES = ["cora_ml", "citeseer"]


def dev():
    return torch.device("cuda:0")


def gpu_ahat(name, mode="sym", **kw):
    import ppnp_b200 as P
    z, adj = load_std(name)
    ip = torch.from_numpy(z["adj_indptr"]).to(dev())
    idx = torch.from_numpy(z["adj_indices"]).to(dev())
    return P.csr_normalize(ip, idx, None, mode, want_val64=True, **kw), adj


# ------------------------------------------------------------------ (1) CSR build + normalisation
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["sym", "rw"])
def test_csr_normalize_bit_exact(name, mode):
    ahat, adj = gpu_ahat(name, mode)
    g = load_golden(name)
    assert np.array_equal(ahat.indptr.cpu().numpy(), g[f"ahat_{mode}_indptr"])
    assert np.array_equal(ahat.indices.cpu().numpy(), g[f"ahat_{mode}_indices"])
    assert np.array_equal(ahat.deg.cpu().numpy(), np.diff(adj.indptr) + 1.0)
    assert np.array_equal(ahat.val64.cpu().numpy(), g[f"ahat_{mode}_data"])          # fp64 bit-exact
    v32 = ahat.val32.cpu().numpy()
    assert np.array_equal(v32, g[f"ahat_{mode}_data"].astype(np.float32))             # rounded once
    assert np.abs(v32 / g[f"ahat_{mode}_data"] - 1).max() <= 2e-7                     # north_star bound


def test_csr_normalize_weighted_with_diagonal_and_empty_rows():
    import ppnp_b200 as P
    rng = np.random.RandomState(0)
    n = 300
    d = (rng.rand(n, n) < 0.03) * rng.rand(n, n)
    d = (d + d.T).astype(np.float32)
    d[5, :] = 0; d[:, 5] = 0            # an isolated node: only the self loop remains
    adj = sp.csr_matrix(d); adj.sort_indices()
    ref = oracle.calc_A_hat(adj, "sym")
    out = P.csr_normalize(torch.from_numpy(adj.indptr).to(dev()), torch.from_numpy(adj.indices).to(dev()),
                          torch.from_numpy(adj.data).to(dev()), "sym", want_val64=True)
    assert np.array_equal(out.indptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(out.indices.cpu().numpy(), ref.indices)
    assert np.array_equal(out.val64.cpu().numpy(), ref.data)


def test_csr_normalize_rmat_matches_c_oracle():
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(20000, 300000, 15, seed=0)
    oip, oidx, oval, odeg = oracle.c_a_hat(ip, idx, None, "sym")
    out = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()), None, "sym", want_val64=True)
    assert np.array_equal(out.indptr.cpu().numpy(), oip)
    assert np.array_equal(out.indices.cpu().numpy(), oidx)
    assert np.array_equal(out.deg.cpu().numpy(), odeg)
    assert np.array_equal(out.val64.cpu().numpy(), oval)


def test_rmat_device_generator_matches_host():
    from ppnp_b200.synth import rmat_adjacency
    ip, idx = oracle.rmat_graph(20000, 300000, 15, seed=3)
    dip, didx = rmat_adjacency(20000, 300000, 15, seed=3, device=dev())
    assert np.array_equal(dip.cpu().numpy(), ip) and np.array_equal(didx.cpu().numpy(), idx)


# ------------------------------------------------------------------------------ (2) APPNP
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("use_vals", [False, True])
@pytest.mark.parametrize("order", ["natural", "degree"])
@pytest.mark.parametrize("per_step", [False, True], ids=["one-launch", "per-step"])
def test_appnp_matches_restatement(name, use_vals, order, per_step):
    """per_step=False: graphs this small run all K steps in the cooperative kernel (stored values in every step, so
    ``use_vals`` changes nothing there); per_step=True: the K per-step launches with the Z2Y / Y / Y2Z epilogues of the
    value-free form (or the plain epilogue with stored values) on the same golden graphs."""
    import ppnp_b200 as P
    ahat, adj = gpu_ahat(name)
    g = load_golden(name)
    graph = P.PropagationGraph(ahat, chunk_edges=128, order=order)
    H = torch.from_numpy(g["H"]).to(dev())
    Zt = P.appnp_propagate(graph, H, K=10, alpha=0.1, use_vals=use_vals, per_step=per_step)
    Z = Zt.cpu().numpy()
    assert relerr(Z, g["appnp_K10"]) < 1e-5                      # north_star fp32 tolerance
    assert (Z.argmax(1) == g["appnp_K10"].argmax(1)).all()       # identical argmax on frozen H
    if per_step:
        one = P.appnp_propagate(graph, H, K=10, alpha=0.1, use_vals=use_vals)
        assert not torch.equal(one, Zt) or use_vals              # really another path (other order of additions) ...
        assert float((one - Zt).norm() / one.norm()) < 1e-6      # ... with the same answer


@pytest.mark.parametrize("F", [1, 3, 7, 16, 20, 64, 100, 256])
@pytest.mark.parametrize("K", [1, 2, 5])
def test_appnp_feature_widths_and_short_K(F, K):
    import ppnp_b200 as P
    ahat, adj = gpu_ahat("citeseer")
    A = oracle.calc_A_hat(adj, "sym")
    Hn = np.random.RandomState(F * 10 + K).randn(adj.shape[0], F).astype(np.float32)
    graph = P.PropagationGraph(ahat, chunk_edges=256)
    for use_vals in (False, True):
        Z = P.appnp_propagate(graph, torch.from_numpy(Hn).to(dev()), K=K, alpha=0.15, use_vals=use_vals).cpu().numpy()
        assert relerr(Z, oracle.appnp(A, Hn.astype(np.float64), 0.15, K)) < 1e-5


def test_appnp_rw_mode():
    import ppnp_b200 as P
    ahat, adj = gpu_ahat("citeseer", "rw")
    A = oracle.calc_A_hat(adj, "rw")
    Hn = np.random.RandomState(2).randn(adj.shape[0], 6).astype(np.float32)
    graph = P.PropagationGraph(ahat, chunk_edges=128)
    for use_vals in (False, True):
        Z = P.appnp_propagate(graph, torch.from_numpy(Hn).to(dev()), K=10, alpha=0.1, use_vals=use_vals).cpu().numpy()
        assert relerr(Z, oracle.appnp(A, Hn.astype(np.float64), 0.1, 10)) < 1e-5


def test_appnp_limit_is_exact_ppnp_kat1():
    import ppnp_b200 as P
    ahat, adj = gpu_ahat("cora_ml")
    g = load_golden("cora_ml")
    graph = P.PropagationGraph(ahat, chunk_edges=128)
    H = torch.from_numpy(g["H"]).to(dev())
    Z = P.appnp_propagate(graph, H, K=200, alpha=0.1).cpu().numpy()
    assert relerr(Z, g["logits_full"]) < 5e-6        # reference ppr @ H (model.py:65), fp32


def test_appnp_backward_is_the_same_operator_kat3():
    import ppnp_b200 as P
    ahat, adj = gpu_ahat("cora_ml")
    g = load_golden("cora_ml")
    A = oracle.calc_A_hat(adj, "sym")
    graph = P.PropagationGraph(ahat, chunk_edges=128)
    H = torch.from_numpy(g["H"]).to(dev()).requires_grad_(True)
    Gn = np.random.RandomState(3).randn(*g["H"].shape).astype(np.float32)
    Z = P.appnp(H, graph, K=10, alpha=0.1)
    Z.backward(torch.from_numpy(Gn).to(dev()))
    assert relerr(H.grad.cpu().numpy(), oracle.appnp(A, Gn.astype(np.float64), 0.1, 10)) < 1e-5
    # adjointness <P H, G> = <H, P G>
    lhs = float((Z.detach().double() * torch.from_numpy(Gn).to(dev()).double()).sum())
    rhs = float((H.detach().double() * H.grad.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))


def test_appnp_rmat_skewed_degrees_vs_c_oracle():
    """Hub rows (split over many chunks), isolated rows (self loop only), F = 64 and 16."""
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(50000, 1200000, 16, seed=0)
    oip, oidx, oval, odeg = oracle.c_a_hat(ip, idx, None, "sym")
    assert (np.diff(ip) == 0).any() and np.diff(ip).max() > 2000
    ahat = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()))
    for order in ("natural", "degree"):
        graph = P.PropagationGraph(ahat, chunk_edges=256, order=order)
        assert graph.plan.n_fix > 0
        for F in (64, 16):
            Hn = np.random.RandomState(F).randn(50000, F).astype(np.float32)
            ref = oracle.c_appnp_f64(oip, oidx, oval, Hn.astype(np.float64), 10, 0.1)
            for use_vals in (False, True):
                Z = P.appnp_propagate(graph, torch.from_numpy(Hn).to(dev()), K=10, alpha=0.1, use_vals=use_vals).cpu().numpy()
                assert relerr(Z, ref) < 1e-5, (order, F, use_vals)


def test_appnp_is_deterministic():
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(30000, 600000, 15, seed=1)
    ahat = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()))
    graph = P.PropagationGraph(ahat)
    H = torch.randn(30000, 64, device=dev())
    a = P.appnp_propagate(graph, H, 10, 0.1)
    b = P.appnp_propagate(graph, H, 10, 0.1)
    assert torch.equal(a, b)


def test_appnp_full_size_properties():
    """BASELINE config 4 size (2 M nodes / ~50 M edges, F = 64): size-independent properties --
    linearity, adjointness, constant-vector fixed point of the 'rw' form is checked through 'sym':
    A_hat (D^1/2 1) = D^1/2 1, so Z = D^1/2 c is a fixed point of the iteration."""
    import ppnp_b200 as P
    from ppnp_b200.synth import rmat_adjacency
    n = 2_000_000
    ip, idx = rmat_adjacency(n, 26_400_000, 21, seed=0, device=dev())
    ahat = P.csr_normalize(ip, idx)
    assert ahat.nnz == int(ip[-1]) + n
    graph = P.PropagationGraph(ahat, chunk_edges=256)
    g = torch.Generator(device=dev()).manual_seed(1)
    H1 = torch.randn(n, 64, device=dev(), generator=g)
    H2 = torch.randn(n, 64, device=dev(), generator=g)
    Z1 = P.appnp_propagate(graph, H1, 10, 0.1)
    Z2 = P.appnp_propagate(graph, H2, 10, 0.1)
    Z12 = P.appnp_propagate(graph, H1 + 2 * H2, 10, 0.1)
    assert float((Z12 - (Z1 + 2 * Z2)).norm() / Z12.norm()) < 1e-5          # linearity
    lhs = float((Z1.double() * H2.double()).sum()); rhs = float((H1.double() * Z2.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), abs(rhs), 1.0)              # adjointness (symmetric A_hat)
    fp = torch.sqrt(ahat.deg).to(torch.float32)[:, None].expand(n, 64).contiguous()
    Zf = P.appnp_propagate(graph, fp, 10, 0.1)
    assert float((Zf - fp).abs().max() / fp.abs().max()) < 1e-5              # fixed point
    Zv = P.appnp_propagate(graph, H1, 10, 0.1, use_vals=True)
    assert float((Zv - Z1).norm() / Z1.norm()) < 1e-5                        # stored values == value-free


# ------------------------------------------------------------------------------ (3) exact PPNP
@pytest.mark.parametrize("name", NAMES)
def test_ppr_dense_matches_reference_inverse(name):
    import ppnp_b200 as P
    ahat, adj = gpu_ahat(name)
    g = load_golden(name)
    Pi = P.ppr_dense(ahat, 0.1, tol=1e-7)
    rows = Pi[torch.from_numpy(g["ppr_rows_idx"]).to(dev())].cpu().numpy()
    assert relerr(rows, g["ppr_rows"]) < 1e-5
    assert relerr(torch.diagonal(Pi).cpu().numpy(), g["ppr_diag"]) < 1e-5
    assert relerr(Pi.sum(1).cpu().numpy(), g["ppr_rowsum"]) < 1e-5
    assert abs(float(Pi.double().norm()) / float(g["ppr_fro"]) - 1) < 1e-5
    assert float((Pi - Pi.T).abs().max()) < 1e-6


def test_ppr_dense_small_K_matches_series():
    import ppnp_b200 as P
    ahat, adj = gpu_ahat("citeseer")
    A = oracle.calc_A_hat(adj, "sym").toarray()
    n = adj.shape[0]
    for K in (0, 1, 2, 3):
        Pi = P.ppr_dense(ahat, 0.2, K=K).cpu().numpy()
        Z = np.eye(n)
        for _ in range(K):
            Z = 0.8 * (A @ Z) + 0.2 * np.eye(n)
        assert relerr(Pi, Z) < 1e-6


@pytest.mark.parametrize("name", NAMES)
def test_gather_gemm_f32_matches_reference_forward_and_grad(name):
    import ppnp_b200 as P
    _, adj = load_std(name)
    g = load_golden(name)
    ppr32 = torch.from_numpy(oracle.compute_ppr(adj, 0.1).astype(np.float32)).to(dev())
    H = torch.from_numpy(g["H"]).to(dev()).requires_grad_(True)
    idx = torch.from_numpy(g["idx_train"]).to(dev())
    out = P.ppr_matmul(ppr32, H, idx)
    assert relerr(out.detach().cpu().numpy(), g["logits_train"]) < 1e-5
    assert (out.detach().cpu().numpy().argmax(1) == g["logits_train"].argmax(1)).all()
    out.backward(torch.from_numpy(g["G_train"]).to(dev()))
    assert relerr(H.grad.cpu().numpy(), g["dH_train"]) < 1e-5
    full = P.gather_gemm(ppr32, H.detach(), None)
    assert relerr(full.cpu().numpy(), g["logits_full"]) < 1e-5


@pytest.mark.parametrize("C", [1, 3, 7, 15, 16, 33, 64, 70])
@pytest.mark.parametrize("m", [1, 60, 333])
def test_gather_gemm_f32_shapes(C, m):
    import ppnp_b200 as P
    rng = np.random.RandomState(C + m)
    n = 1237
    Pi = rng.rand(n, n).astype(np.float32)
    H = rng.randn(n, C).astype(np.float32)
    idx = rng.choice(n, m, replace=True)
    out = P.gather_gemm(torch.from_numpy(Pi).to(dev()), torch.from_numpy(H).to(dev()), torch.from_numpy(idx).to(dev()))
    assert relerr(out.cpu().numpy(), Pi[idx].astype(np.float64) @ H) < 1e-5
    G = rng.randn(m, C).astype(np.float32)
    outT = P.gather_gemm(torch.from_numpy(Pi).to(dev()), torch.from_numpy(G).to(dev()), torch.from_numpy(idx).to(dev()), transpose=True)
    assert relerr(outT.cpu().numpy(), Pi[idx].astype(np.float64).T @ G) < 1e-5


# --------------------------------------------------------------------------- (4) batch-main path
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("k", [1, 32, 128])
def test_topk_threshold_and_mask_bit_exact_on_reference_ppr(name, k):
    import ppnp_b200 as P
    _, adj = load_std(name)
    ppr32 = oracle.compute_ppr(adj, 0.1).astype(np.float32)
    t = torch.from_numpy(ppr32).to(dev())
    th = P.topk_thresh(t, k).cpu().numpy()
    assert np.array_equal(th, oracle.topk_thresh(ppr32, k))                        # selection is exact
    assert np.array_equal(th, torch.from_numpy(ppr32).topk(k, -1).values[:, -1].numpy())  # the literal line 115
    P.topk_sparsify_(t, k)
    assert np.array_equal(t.cpu().numpy(), oracle.topk_sparsify(ppr32, k))         # the literal line 116


def test_topk_handles_ties_negatives_and_k_equal_n():
    import ppnp_b200 as P
    rng = np.random.RandomState(0)
    a = rng.randint(-3, 4, size=(97, 97)).astype(np.float32)       # many ties, negatives, zeros
    for k in (1, 5, 97):
        th = P.topk_thresh(torch.from_numpy(a).to(dev()), k).cpu().numpy()
        assert np.array_equal(th, torch.from_numpy(a).topk(k, -1).values[:, -1].numpy())
    with pytest.raises(RuntimeError):
        P.topk_thresh(torch.from_numpy(a).to(dev()), 98)


@pytest.mark.parametrize("name", NAMES)
def test_batch_step_matches_reference_lines(name):
    import ppnp_b200 as P
    _, adj = load_std(name)
    g = load_golden(name)
    k = int(g["topk_k"])
    ppr32 = oracle.compute_ppr(adj, 0.1).astype(np.float32)
    dense = oracle.topk_sparsify(ppr32, k)
    t = torch.from_numpy(ppr32).to(dev())
    P.topk_sparsify_(t, k)
    spp = P.dense_to_sparse_ppr(t)
    ref_csr = sp.csr_matrix(dense)
    assert np.array_equal(spp.indptr.cpu().numpy(), ref_csr.indptr)
    assert np.array_equal(spp.indices.cpu().numpy(), ref_csr.indices)
    assert np.array_equal(spp.val.cpu().numpy(), ref_csr.data)
    for B, seed in ((1, 0), (32, 3), (140, 4), (1024, 5)):
        idx_b = np.sort(np.random.RandomState(seed).choice(adj.shape[0], min(B, adj.shape[0]), replace=False))
        logits_ref, sel_ref = oracle.batch_step(dense, idx_b, g["H"].astype(np.float64))
        sel = P.batch_support(spp, torch.from_numpy(idx_b).to(dev()))
        assert np.array_equal(sel.cpu().numpy(), sel_ref)
        Hsub = torch.from_numpy(g["H"]).to(dev())[sel].requires_grad_(True)
        out = P.batch_propagate(spp, torch.from_numpy(idx_b).to(dev()), sel, Hsub)
        assert relerr(out.detach().cpu().numpy(), logits_ref) < 1e-5
        Gn = np.random.RandomState(seed + 9).randn(*logits_ref.shape).astype(np.float32)
        out.backward(torch.from_numpy(Gn).to(dev()))
        dref = dense[idx_b][:, sel_ref].astype(np.float64).T @ Gn
        assert relerr(Hsub.grad.cpu().numpy(), dref) < 1e-5
        # the one-launch form: the same index tensor for both calls -> the column map made with the mask is used
        ib = torch.from_numpy(idx_b).to(dev())
        sel1 = P.batch_support(spp, ib)
        colmap, _ = sel1._ppnp_colmap
        want = np.where(sel_ref, np.cumsum(sel_ref) - 1, -1)
        assert np.array_equal(sel1.cpu().numpy(), sel_ref) and np.array_equal(colmap.cpu().numpy(), want)
        H1 = torch.from_numpy(g["H"]).to(dev())[sel1].requires_grad_(True)
        out1 = P.batch_propagate(spp, ib, sel1, H1)
        assert torch.equal(out1, out)
        out1.backward(torch.from_numpy(Gn).to(dev()))
        assert relerr(H1.grad.cpu().numpy(), dref) < 1e-5
        # arbitrary masks are valid, like the reference's ppr_sub[:, mask] (columns outside the mask are dropped)
        rs = np.random.RandomState(seed + 1)
        for mask in (sel_ref | (rs.rand(len(sel_ref)) < 0.3), sel_ref & (rs.rand(len(sel_ref)) < 0.6)):
            if not mask.any():
                continue
            tm = torch.from_numpy(mask).to(dev())
            outm = P.batch_propagate(spp, ib, tm, torch.from_numpy(g["H"]).to(dev())[tm])
            refm = dense[idx_b][:, mask].astype(np.float64) @ g["H"].astype(np.float64)[mask]
            assert relerr(outm.cpu().numpy(), refm) < 1e-5


# ------------------------------------------------------------------ persistent K-step kernel
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("F", [3, 7, 16, 64])
def test_persistent_kernel_matches_restatement(name, F):
    """All K steps in one cooperative launch (grid barriers) == the per-step launches == the oracle."""
    import ppnp_b200 as P
    ahat, adj = gpu_ahat(name)
    A = oracle.calc_A_hat(adj, "sym")
    Hn = np.random.RandomState(F).randn(adj.shape[0], F).astype(np.float32)
    H = torch.from_numpy(Hn).to(dev())
    for order in ("natural", "degree"):
        graph = P.PropagationGraph(ahat, chunk_edges=128, order=order)
        for K in (1, 2, 10, 11):
            ref = oracle.appnp(A, Hn.astype(np.float64), 0.1, K)
            Zp = P.appnp_propagate_persistent(graph, H, K, 0.1)
            assert relerr(Zp.cpu().numpy(), ref) < 1e-5, (order, K)
            Zp2 = P.appnp_propagate_persistent(graph, H, K, 0.1)
            assert torch.equal(Zp, Zp2)                                  # deterministic
        # the public entry point picks the persistent kernel for a graph this small
        Za = P.appnp_propagate(graph, H, 10, 0.1)
        assert torch.equal(Za, P.appnp_propagate_persistent(graph, H, 10, 0.1))


def test_persistent_kernel_on_hub_rows():
    """Rows split over many chunks need the in-kernel fix-up phase between two grid barriers."""
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(30000, 500000, 15, seed=2)
    oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
    ahat = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()))
    graph = P.PropagationGraph(ahat, chunk_edges=128, order="degree")
    assert graph.plan.n_fix > 0 and graph.plan.n_chunks <= 4096 * 2
    Hn = np.random.RandomState(0).randn(30000, 16).astype(np.float32)
    ref = oracle.c_appnp_f64(oip, oidx, oval, Hn.astype(np.float64), 10, 0.1)
    Z = P.appnp_propagate_persistent(graph, torch.from_numpy(Hn).to(dev()), 10, 0.1)
    assert relerr(Z.cpu().numpy(), ref) < 1e-5


# ------------------------------------------------------------------ edge cases and error behaviour
def _tiny_graph(n_edges_dir, n):
    import scipy.sparse as sp_
    rng = np.random.RandomState(n)
    r = rng.randint(0, n, n_edges_dir); c = rng.randint(0, n, n_edges_dir)
    keep = r != c
    a = sp_.csr_matrix((np.ones(keep.sum(), np.float32), (r[keep], c[keep])), shape=(n, n))
    a = ((a + a.T) > 0).astype(np.float32).tocsr()
    a.sort_indices()
    return a


@pytest.mark.parametrize("n,m", [(1, 0), (2, 1), (3, 0), (33, 40), (257, 300)])
def test_tiny_and_edgeless_graphs(n, m):
    """n = 1, isolated nodes only, graphs far smaller than one chunk: every row is just its self loop or
    a handful of edges, the stream is almost all padding."""
    import ppnp_b200 as P
    adj = _tiny_graph(max(m, 1), n) if m else __import__("scipy.sparse").sparse.csr_matrix((n, n), dtype=np.float32)
    A = oracle.calc_A_hat(adj, "sym")
    ahat = P.csr_normalize(torch.from_numpy(adj.indptr.astype(np.int32)).to(dev()),
                           torch.from_numpy(adj.indices.astype(np.int32)).to(dev()), want_val64=True)
    assert np.array_equal(ahat.indptr.cpu().numpy(), A.indptr) and np.array_equal(ahat.indices.cpu().numpy(), A.indices)
    assert np.array_equal(ahat.val64.cpu().numpy(), A.data)
    Hn = np.random.RandomState(1).randn(n, 5).astype(np.float32)
    for order in ("natural", "degree"):
        graph = P.PropagationGraph(ahat, chunk_edges=128, order=order)
        for K in (1, 3):
            for use_vals in (False, True):
                Z = P.appnp_propagate(graph, torch.from_numpy(Hn).to(dev()), K, 0.1, use_vals=use_vals).cpu().numpy()
                assert relerr(Z, oracle.appnp(A, Hn.astype(np.float64), 0.1, K)) < 1e-5
    Pi = P.ppr_dense(ahat, 0.1, tol=1e-7).cpu().numpy()
    assert relerr(Pi, oracle.compute_ppr(adj, 0.1)) < 1e-5


def test_argument_errors_are_reported_not_crashes():
    import ppnp_b200 as P
    from ppnp_b200 import _lib
    ahat, adj = gpu_ahat("citeseer")
    graph = P.PropagationGraph(ahat, chunk_edges=128)
    n = adj.shape[0]
    H = torch.randn(n, 4, device=dev())
    with pytest.raises(ValueError):
        P.appnp_propagate(graph, torch.randn(n + 1, 4, device=dev()), 3, 0.1)          # wrong number of rows
    with pytest.raises(ValueError):
        P.appnp_propagate(graph, H.double(), 3, 0.1)                                    # wrong dtype
    with pytest.raises(RuntimeError, match="alias"):
        P.spmm_step(graph, H, H, 0.1, out=H)                                            # in-place step
    with pytest.raises(RuntimeError, match="epilogue"):
        P.spmm_step(graph, H, H, 0.1, epi=9)
    assert torch.equal(P.appnp_propagate(graph, H, 0, 0.1), H)                          # K = 0: Z_0 = H
    with pytest.raises(ValueError):
        P.PropagationGraph(ahat, chunk_edges=100)                                       # not a multiple of 128
    Pi = torch.rand(50, 50, device=dev())
    with pytest.raises(ValueError):
        P.gather_gemm(Pi, torch.randn(49, 3, device=dev()))                             # inner dimensions differ
    with pytest.raises(RuntimeError, match="multiple of 8"):
        P.gather_gemm_bf16(Pi.to(torch.bfloat16), torch.randn(50, 3, device=dev()))    # unpadded bf16 rows
    assert P.gather_gemm(Pi, torch.randn(50, 3, device=dev()), torch.zeros(0, dtype=torch.int64, device=dev())).shape == (0, 3)
    # the message of the last failure is retrievable through the C ABI
    assert b"multiple of 8" in _lib.load().ppnp_last_error()


def test_appnp_weighted_adjacency_with_diagonal_never_takes_the_value_free_path():
    """the value-free iteration reads a row's degree off its edge count, which is D only for unit weights and
    an empty diagonal.  A weighted graph with self loops, large enough for the per-step launches (> 4096 chunks), must give
    the oracle's result whatever `use_vals` says, and the value-free single step must refuse it."""
    import ppnp_b200 as P
    from ppnp_b200 import _lib
    ip, idx = oracle.rmat_graph(30000, 600000, 15, seed=9)
    n = len(ip) - 1
    rng = np.random.RandomState(0)
    adj = sp.csr_matrix((rng.rand(len(idx)).astype(np.float32) + 0.5, idx, ip), shape=(n, n))
    adj = (adj + adj.T + sp.diags((rng.rand(n) < 0.3).astype(np.float32) * 2.0)).tocsr().astype(np.float32)
    adj.sort_indices()
    ahat = P.csr_normalize(torch.from_numpy(adj.indptr).to(dev()), torch.from_numpy(adj.indices).to(dev()),
                           torch.from_numpy(adj.data).to(dev()), "sym")
    assert not ahat.unit_weights
    g = P.PropagationGraph(ahat, chunk_edges=128, order="degree")
    assert g.plan.n_chunks > 4096
    H = rng.randn(n, 16).astype(np.float32)
    ref = oracle.appnp(oracle.calc_A_hat(adj, "sym"), H.astype(np.float64), 0.1, 10)
    for uv in (False, True):
        z = P.appnp_propagate(g, torch.from_numpy(H).to(dev()), 10, 0.1, use_vals=uv).cpu().numpy()
        assert relerr(z, ref) < 1e-5
    with pytest.raises(ValueError):
        P.spmm_step(g, torch.from_numpy(H).to(dev()), torch.from_numpy(H).to(dev()), 0.1, _lib.EPI_Y, False)
    with pytest.raises(ValueError):
        P.PropagationGraph(ahat, keep_vals=False)
    # an all-ones adjacency WITH stored diagonal entries is not "unit" either
    adj1 = (sp.csr_matrix((np.ones(len(idx), np.float32), idx, ip), shape=(n, n)) + sp.eye(n, format="csr", dtype=np.float32)).tocsr()
    adj1.sort_indices()
    a1 = P.csr_normalize(torch.from_numpy(adj1.indptr).to(dev()), torch.from_numpy(adj1.indices).to(dev()), None, "sym")
    assert not a1.unit_weights
    z = P.appnp_propagate(P.PropagationGraph(a1, chunk_edges=128), torch.from_numpy(H).to(dev()), 10, 0.1).cpu().numpy()
    assert relerr(z, oracle.appnp(oracle.calc_A_hat(adj1, "sym"), H.astype(np.float64), 0.1, 10)) < 1e-5
