"""The shared-memory-resident hub-row step (csrc/appnp_tiled.cu) on a GPU, through the C ABI, against the fp64 C
oracle and against the row-major kernel.  SURVEY.md section 8 row "APPNP K-step" (north_star subsystem 2)."""
import numpy as np
import pytest
import torch

from util import load_std, oracle, relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _graph(n, raw, scale, seed, **kw):
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(n, raw, scale, seed=seed)
    ahat = P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()))
    return ip, idx, ahat, P.PropagationGraph(ahat, chunk_edges=128, order="degree", **kw)


TILED = [
    dict(slice_width=64, n_ctas=20, warps_per_cta=8, slot_rows=64, min_hub_degree=8, fine_cols=64, coarse_edges=2048),
    dict(slice_width=32, n_ctas=12, warps_per_cta=16, slot_rows=128, min_hub_degree=4, fine_cols=128, coarse_edges=1024),
    dict(slice_width=16, n_ctas=7, warps_per_cta=4, slot_rows=200, min_hub_degree=16, fine_cols=32, coarse_edges=512, slack=0),
    dict(slice_width=64, min_hub_degree=32),      # the defaults the bench uses: one CTA per SM
]


@pytest.mark.parametrize("tiled", TILED)
@pytest.mark.parametrize("F", [64, 128])
def test_tiled_step_matches_oracle(tiled, F):
    import ppnp_b200 as P
    from ppnp_b200 import _lib
    ip, idx, ahat, g = _graph(30000, 600000, 15, 4, tiled=tiled)
    oip, oidx, oval, odeg = oracle.c_a_hat(ip, idx, None, "sym")
    n = len(ip) - 1
    rng = np.random.RandomState(F)
    Zin, T = rng.randn(n, F).astype(np.float32), rng.randn(n, F).astype(np.float32)
    tl = g.tiled_for(F)
    assert tl is not None and tl[0].stats["hub_rows"] > 0
    import scipy.sparse as sp
    A = sp.csr_matrix((oval, oidx, oip), shape=(n, n))
    ones = sp.csr_matrix((np.ones_like(oval), oidx, oip), shape=(n, n))
    for epi, uv in [(_lib.EPI_PLAIN, True), (_lib.EPI_Z2Y, True), (_lib.EPI_Y, False), (_lib.EPI_Y2Z, False), (_lib.EPI_RW, False)]:
        out = P.spmm_step(g, torch.from_numpy(Zin).to(dev()), torch.from_numpy(T).to(dev()), 0.1, epi, uv).cpu().numpy()
        acc = (A if uv else ones) @ Zin.astype(np.float64)
        d = odeg
        a, b = {_lib.EPI_PLAIN: (0.9 + 0 * d, 0.1 + 0 * d), _lib.EPI_Z2Y: (0.9 / np.sqrt(d), 0.1 / np.sqrt(d)),
                _lib.EPI_Y: (0.9 / d, 0.1 / np.sqrt(d)), _lib.EPI_Y2Z: (0.9 / np.sqrt(d), 0.1 + 0 * d),
                _lib.EPI_RW: (0.9 / d, 0.1 + 0 * d)}[epi]
        ref = a[:, None] * acc + b[:, None] * T
        assert relerr(out, ref) < 2e-6, (epi, uv, relerr(out, ref))
        hub = tl[0].hub_rows.cpu().numpy()
        assert relerr(out[hub], ref[hub]) < 2e-6


@pytest.mark.parametrize("tiled", TILED)
def test_tiled_appnp_matches_oracle_and_row_major(tiled):
    import ppnp_b200 as P
    ip, idx, ahat, g = _graph(30000, 600000, 15, 5, tiled=tiled)
    g0 = P.PropagationGraph(ahat, chunk_edges=128, order="degree")
    oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
    n, F, K = len(ip) - 1, 64, 10
    H = np.random.RandomState(1).randn(n, F).astype(np.float32)
    ref = oracle.c_appnp_f64(oip, oidx, oval, H.astype(np.float64), K, 0.1)
    Hd = torch.from_numpy(H).to(dev())
    for uv in (False, True):
        z = P.appnp_propagate(g, Hd, K, 0.1, use_vals=uv).cpu().numpy()
        z0 = P.appnp_propagate(g0, Hd, K, 0.1, use_vals=uv).cpu().numpy()
        assert relerr(z, ref) < 1e-5 and relerr(z, z0) < 1e-5
        assert (z.argmax(1) == ref.argmax(1)).mean() > 0.9999
    # adjointness (KAT-3): <P(H), G> == <H, P(G)>
    G = torch.from_numpy(np.random.RandomState(2).randn(n, F).astype(np.float32)).to(dev())
    lhs = (P.appnp_propagate(g, Hd, K, 0.1).double() * G.double()).sum().item()
    rhs = (Hd.double() * P.appnp_propagate(g, G, K, 0.1).double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_tiled_autograd_on_cora_ml():
    """Every row a hub row (no row-major part), through the differentiable op."""
    import ppnp_b200 as P
    import scipy.sparse as sp
    z, adj = load_std("cora_ml")
    ahat = P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev()), torch.from_numpy(z["adj_indices"]).to(dev()))
    g = P.PropagationGraph(ahat, chunk_edges=128, tiled=dict(slice_width=16, n_ctas=16, warps_per_cta=8, slot_rows=256,
                                                             min_hub_degree=1, fine_cols=64, coarse_edges=256))
    n, F = ahat.n, 16
    assert g.tiled_for(F)[1] is None
    H = torch.randn(n, F, device=dev(), generator=torch.Generator(device=dev()).manual_seed(0)).requires_grad_(True)
    Z = P.appnp(H, g, K=10, alpha=0.1)
    G = torch.randn(n, F, device=dev(), generator=torch.Generator(device=dev()).manual_seed(1))
    Z.backward(G)
    A = oracle.calc_A_hat(adj, "sym")
    Zo = oracle.appnp(A, H.detach().cpu().numpy().astype(np.float64), 0.1, 10)
    dHo = oracle.appnp(A, G.cpu().numpy().astype(np.float64), 0.1, 10)
    assert relerr(Z.detach().cpu().numpy(), Zo) < 1e-5 and relerr(H.grad.cpu().numpy(), dHo) < 1e-5


def test_tiled_rejects_what_it_cannot_run():
    import ppnp_b200 as P
    from ppnp_b200 import _lib
    ip, idx, ahat, g = _graph(5000, 60000, 13, 6, tiled=dict(slice_width=32, n_ctas=4, warps_per_cta=4, slot_rows=64, min_hub_degree=8))
    assert g.tiled_for(48) is None          # not a multiple of the slice width: the row-major kernel runs
    tp, rest, W, _ = g.tiled_for(64)
    lib = _lib.load()
    Z = torch.zeros(ahat.n, 64, device=dev())
    rc = lib.ppnp_spmm_step_tiled(tp.struct(), _lib.ptr(Z), _lib.ptr(Z), _lib.ptr(Z), 64, 64, 32, 0.1, 0, 0, _lib.current_stream())
    assert rc == -1 and b"aliased" in lib.ppnp_last_error()
    rc = lib.ppnp_spmm_step_tiled(tp.struct(), _lib.ptr(Z), _lib.ptr(Z), _lib.ptr(torch.empty_like(Z)), 64, 64, 24, 0.1, 0, 0,
                                  _lib.current_stream())
    assert rc == -1
