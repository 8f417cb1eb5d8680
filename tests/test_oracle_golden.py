"""The oracle (oracle/ppnp_oracle.py + ppnp_oracle.c) against golden vectors produced by the
reference itself (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from util import load_golden, load_std, oracle, relerr

NAMES = ["cora_ml", "citeseer"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["sym", "rw"])
def test_calc_A_hat_numpy(name, mode):
    _, adj = load_std(name)
    g = load_golden(name)
    ah = oracle.calc_A_hat(adj, mode)
    assert np.array_equal(ah.indptr, g[f"ahat_{mode}_indptr"])
    assert np.array_equal(ah.indices, g[f"ahat_{mode}_indices"])
    assert np.array_equal(ah.data, g[f"ahat_{mode}_data"])  # fp64 bit-exact


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["sym", "rw"])
def test_calc_A_hat_c(name, mode):
    z, _ = load_std(name)
    g = load_golden(name)
    ip, idx, val, deg = oracle.c_a_hat(z["adj_indptr"], z["adj_indices"], None, mode)
    assert np.array_equal(ip, g[f"ahat_{mode}_indptr"])
    assert np.array_equal(idx, g[f"ahat_{mode}_indices"])
    assert np.array_equal(val, g[f"ahat_{mode}_data"])
    assert np.array_equal(deg, np.diff(z["adj_indptr"]) + 1.0)


def test_calc_A_hat_weighted_and_diagonal():
    # beyond the fixtures: weights and an existing diagonal entry (adj + I adds 1 to it)
    import scipy.sparse as sp
    rng = np.random.RandomState(0)
    n = 50
    d = (rng.rand(n, n) < 0.1) * rng.rand(n, n)
    d = (d + d.T).astype(np.float32)
    adj = sp.csr_matrix(d)
    adj.sort_indices()
    ref = oracle.calc_A_hat(adj, "sym")
    ip, idx, val, deg = oracle.c_a_hat(adj.indptr, adj.indices, adj.data, "sym")
    assert np.array_equal(ip, ref.indptr) and np.array_equal(idx, ref.indices)
    np.testing.assert_allclose(val, ref.data, rtol=1e-14)


@pytest.mark.parametrize("name", NAMES)
def test_compute_ppr(name):
    _, adj = load_std(name)
    g = load_golden(name)
    ppr = oracle.compute_ppr(adj, float(g["alpha"]))
    np.testing.assert_allclose(ppr[g["ppr_rows_idx"]], g["ppr_rows"], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(np.diag(ppr), g["ppr_diag"], rtol=1e-10)
    np.testing.assert_allclose(ppr.sum(1), g["ppr_rowsum"], rtol=1e-10)
    # symmetric, positive on a connected graph
    assert np.abs(ppr - ppr.T).max() < 1e-14 and ppr.min() > 0


@pytest.mark.parametrize("name", NAMES)
def test_forward_and_grad(name):
    _, adj = load_std(name)
    g = load_golden(name)
    ppr32 = oracle.compute_ppr(adj, float(g["alpha"])).astype(np.float32)
    H = g["H"]
    idx = g["idx_train"]
    assert relerr(oracle.ppnp_forward(ppr32, H, idx), g["logits_train"]) < 1e-6
    assert relerr(oracle.ppnp_forward(ppr32, H), g["logits_full"]) < 1e-6
    assert relerr(oracle.ppnp_forward_grad(ppr32, g["G_train"], idx), g["dH_train"]) < 1e-6


@pytest.mark.parametrize("name", NAMES)
def test_topk_and_batch(name):
    _, adj = load_std(name)
    g = load_golden(name)
    ppr32 = oracle.compute_ppr(adj, float(g["alpha"])).astype(np.float32)
    k = int(g["topk_k"])
    th = oracle.topk_thresh(ppr32, k)
    # the fp32 cast of the LAPACK result may differ in the last bit between runs of different
    # BLAS builds; thresholds are data, compare as floats
    np.testing.assert_allclose(th, g["topk_thresh"], rtol=1e-6)
    sp_ppr = oracle.topk_sparsify(ppr32, k)
    # quirk check: column j keeps >= k entries (ties kept), rows do not
    assert ((sp_ppr > 0).sum(0) >= k).all()
    same_rows = np.mean((sp_ppr > 0).sum(1) == g["topk_row_nnz"])
    assert same_rows > 0.98
    logits, sel = oracle.batch_step(sp_ppr, g["batch_idx"], g["H"])
    assert np.mean(sel == g["batch_sel"]) > 0.995
    if np.array_equal(sel, g["batch_sel"]):
        assert relerr(logits, g["batch_logits"]) < 1e-5


def test_topk_c_matches_numpy():
    rng = np.random.RandomState(1)
    a = rng.rand(64, 64).astype(np.float32)
    a = (a + a.T) / 2
    th = np.empty(64, dtype=np.float32)
    lib = oracle.clib()
    lib.oracle_topk_thresh(64, 64, oracle._ptr(a), 5, oracle._ptr(th))
    assert np.array_equal(th, oracle.topk_thresh(a, 5))
    b = a.copy()
    lib.oracle_topk_mask(64, oracle._ptr(b), oracle._ptr(th))
    assert np.array_equal(b, oracle.topk_sparsify(a, 5))


@pytest.mark.parametrize("name", NAMES)
def test_appnp_restatement(name):
    _, adj = load_std(name)
    g = load_golden(name)
    alpha = float(g["alpha"])
    ah = oracle.calc_A_hat(adj, "sym")
    H = g["H"].astype(np.float64)
    Z = oracle.appnp(ah, H, alpha, 10)
    assert relerr(Z, g["appnp_K10"]) < 1e-14
    # C port, fp64 and fp32
    Zc = oracle.c_appnp_f64(ah.indptr, ah.indices, ah.data, H, 10, alpha)
    assert relerr(Zc, g["appnp_K10"]) < 1e-13
    Zf = oracle.c_appnp_f32(ah.indptr, ah.indices, ah.data, g["H"], 10, alpha)
    assert relerr(Zf, g["appnp_K10"]) < 2e-6
    # KAT-1: K -> inf equals compute_ppr @ H;  KAT-3: adjointness
    ppr = oracle.compute_ppr(adj, alpha)
    assert relerr(oracle.appnp(ah, H, alpha, 300), ppr @ H) < 1e-12
    G = np.random.RandomState(5).randn(*H.shape)
    assert abs(np.sum(oracle.appnp(ah, H, alpha, 10) * G) - np.sum(H * oracle.appnp(ah, G, alpha, 10))) < 1e-9


def test_kat2_identity_gives_ppr():
    _, adj = load_std("citeseer")
    ah = oracle.calc_A_hat(adj, "sym")
    n = adj.shape[0]
    ppr = oracle.compute_ppr(adj, 0.1)
    Z = oracle.appnp(ah, np.eye(n), 0.1, 250)
    assert relerr(Z, ppr) < 1e-10


def test_rmat_generator_is_deterministic_and_symmetric():
    ip, idx = oracle.rmat_graph(5000, 60000, 13, seed=0)
    ip2, idx2 = oracle.rmat_graph(5000, 60000, 13, seed=0)
    assert np.array_equal(ip, ip2) and np.array_equal(idx, idx2)
    import scipy.sparse as sp
    A = sp.csr_matrix((np.ones(len(idx)), idx, ip), shape=(5000, 5000))
    assert (A != A.T).nnz == 0 and A.diagonal().sum() == 0
    assert all(np.all(np.diff(idx[ip[i]:ip[i + 1]]) > 0) for i in range(0, 5000, 97))
