"""The one-lane-group-per-row kernel (csrc/appnp_rows.cu) through the C ABI against the fp64 C oracle, alone and as the
low-degree part next to the edge stream / the tiled hub kernel.  SURVEY.md section 8 row "APPNP K-step"."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from util import load_std, oracle, relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _ahat(n, raw, scale, seed):
    import ppnp_b200 as P
    ip, idx = oracle.rmat_graph(n, raw, scale, seed=seed)
    return ip, idx, P.csr_normalize(torch.from_numpy(ip.astype(np.int32)).to(dev()), torch.from_numpy(idx).to(dev()))


@pytest.mark.parametrize("F", [4, 8, 16, 28, 64, 128, 192])
@pytest.mark.parametrize("rows_below", [1 << 30, 24])
def test_rows_step_matches_oracle(F, rows_below):
    """rows_below = 2^30: every row through the rows kernel; 24: hubs through the stream, the rest through the rows kernel."""
    import ppnp_b200 as P
    from ppnp_b200 import _lib
    ip, idx, ahat = _ahat(20000, 300000, 15, 2)
    g = P.PropagationGraph(ahat, chunk_edges=128, order="degree", rows_below=rows_below)
    assert g.rows_part is not None and (g.plan is None) == (rows_below > 10 ** 6)
    oip, oidx, oval, odeg = oracle.c_a_hat(ip, idx, None, "sym")
    n = len(ip) - 1
    rng = np.random.RandomState(F)
    Zin, T = rng.randn(n, F).astype(np.float32), rng.randn(n, F).astype(np.float32)
    A = sp.csr_matrix((oval, oidx, oip), shape=(n, n))
    ones = sp.csr_matrix((np.ones_like(oval), oidx, oip), shape=(n, n))
    d = odeg
    for epi, uv in [(_lib.EPI_PLAIN, True), (_lib.EPI_Z2Y, True), (_lib.EPI_Y, False), (_lib.EPI_Y2Z, False), (_lib.EPI_RW, False)]:
        out = P.spmm_step(g, torch.from_numpy(Zin).to(dev()), torch.from_numpy(T).to(dev()), 0.1, epi, uv).cpu().numpy()
        acc = (A if uv else ones) @ Zin.astype(np.float64)
        a, b = {_lib.EPI_PLAIN: (0.9 + 0 * d, 0.1 + 0 * d), _lib.EPI_Z2Y: (0.9 / np.sqrt(d), 0.1 / np.sqrt(d)),
                _lib.EPI_Y: (0.9 / d, 0.1 / np.sqrt(d)), _lib.EPI_Y2Z: (0.9 / np.sqrt(d), 0.1 + 0 * d),
                _lib.EPI_RW: (0.9 / d, 0.1 + 0 * d)}[epi]
        assert relerr(out, a[:, None] * acc + b[:, None] * T) < 2e-6, (epi, uv)


@pytest.mark.parametrize("kw", [dict(rows_below=1 << 30), dict(rows_below=32), dict(rows_below=32, idx16=True),
                                dict(tiled=dict(slice_width=64, n_ctas=20, warps_per_cta=8, slot_rows=64, min_hub_degree=8, rest="rows")),
                                dict(tiled=dict(slice_width=32, rest="rows"))])
def test_rows_appnp_matches_oracle_and_is_deterministic(kw):
    import ppnp_b200 as P
    ip, idx, ahat = _ahat(30000, 600000, 15, 5)
    g = P.PropagationGraph(ahat, chunk_edges=128, order="degree", **kw)
    oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
    n, F, K = len(ip) - 1, 64, 10
    H = np.random.RandomState(1).randn(n, F).astype(np.float32)
    ref = oracle.c_appnp_f64(oip, oidx, oval, H.astype(np.float64), K, 0.1)
    Hd = torch.from_numpy(H).to(dev())
    for uv in (False, True):
        z = P.appnp_propagate(g, Hd, K, 0.1, use_vals=uv)
        assert relerr(z.cpu().numpy(), ref) < 1e-5
        if "tiled" not in kw:
            assert torch.equal(z, P.appnp_propagate(g, Hd, K, 0.1, use_vals=uv))       # bit-reproducible
    G = torch.from_numpy(np.random.RandomState(2).randn(n, F).astype(np.float32)).to(dev())
    lhs = (P.appnp_propagate(g, Hd, K, 0.1).double() * G.double()).sum().item()
    rhs = (Hd.double() * P.appnp_propagate(g, G, K, 0.1).double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_rows_on_the_reference_graphs_with_autograd():
    import ppnp_b200 as P
    for name in ("cora_ml", "citeseer"):
        z, adj = load_std(name)
        ahat = P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev()), torch.from_numpy(z["adj_indices"]).to(dev()))
        g = P.PropagationGraph(ahat, chunk_edges=128, order="degree", rows_below=1 << 30)
        n, F = ahat.n, 7
        H = torch.randn(n, 8, device=dev(), generator=torch.Generator(device=dev()).manual_seed(0)).requires_grad_(True)
        Z = P.appnp(H, g, K=10, alpha=0.1)
        G = torch.randn(n, 8, device=dev(), generator=torch.Generator(device=dev()).manual_seed(1))
        Z.backward(G)
        A = oracle.calc_A_hat(adj, "sym")
        assert relerr(Z.detach().cpu().numpy(), oracle.appnp(A, H.detach().cpu().numpy().astype(np.float64), 0.1, 10)) < 1e-5
        assert relerr(H.grad.cpu().numpy(), oracle.appnp(A, G.cpu().numpy().astype(np.float64), 0.1, 10)) < 1e-5


def test_rows_rejects_bad_arguments():
    from ppnp_b200 import _lib
    lib = _lib.load()
    z = torch.zeros(8, 8, device=dev())
    ip = torch.arange(9, dtype=torch.int32, device=dev())
    idx = torch.arange(8, dtype=torch.int32, device=dev())
    rc = lib.ppnp_spmm_step_rows(_lib.ptr(ip), _lib.ptr(idx), None, _lib.ptr(idx), 8, 8, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), 8, 8, 0.1, 0, 0,
                                 None, None, None, None, 0, _lib.current_stream())
    assert rc == -1 and b"aliased" in lib.ppnp_last_error()
    out = torch.empty_like(z)
    rc = lib.ppnp_spmm_step_rows(_lib.ptr(ip), _lib.ptr(idx), None, _lib.ptr(idx), 8, 8, _lib.ptr(z), _lib.ptr(z), _lib.ptr(out), 8, 6, 0.1, 0, 0,
                                 None, None, None, None, 0, _lib.current_stream())
    assert rc == -1
    rc = lib.ppnp_spmm_step_rows(_lib.ptr(ip), _lib.ptr(idx), None, _lib.ptr(idx), 8, 8, _lib.ptr(z), _lib.ptr(z), _lib.ptr(out), 8, 8, 0.1, 0, 1,
                                 None, None, None, None, 0, _lib.current_stream())
    assert rc == -1 and b"stored values" in lib.ppnp_last_error()
