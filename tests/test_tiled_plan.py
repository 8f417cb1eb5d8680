"""Host logic of the shared-memory-resident hub-row step (ppnp_b200/tiled.py) without a GPU: the plan is walked
edge by edge in numpy (tests/util.py walk_tiled, which also asserts the invariants csrc/appnp_tiled.cu relies on)
and, together with the row-major stream over the remaining rows, must reproduce one propagation step."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from util import epi_coef, load_std, oracle, walk_stream, walk_tiled
from ppnp_b200.tiled import build_tiled_plan, choose_windows


def _ahat(ip, idx):
    n = len(ip) - 1
    A = sp.csr_matrix((np.ones(len(idx)), idx, ip), shape=(n, n)) + sp.eye(n, format="csr")
    A.sort_indices()
    d = np.asarray(A.sum(1)).ravel()
    vals = (A.data / np.sqrt(np.repeat(d, np.diff(A.indptr))) / np.sqrt(d[A.indices])).astype(np.float32)
    return A, d, vals


def _check(A, d, vals, tp, F=5, seed=0):
    n = A.shape[0]
    rng = np.random.RandomState(seed)
    Z, T = rng.randn(n, F), rng.randn(n, F)
    for epi, uv in [(0, True), (1, True), (2, False), (3, False), (4, False)]:
        out = walk_tiled(tp, Z, T, 0.1, epi, uv)
        if tp.rest is not None:
            rest = walk_stream(tp.rest, Z, T, 0.1, epi, uv, rows=tp.rest.order.numpy())
            out = np.where(np.isnan(out), rest, out)
        assert not np.isnan(out).any()
        M = sp.csr_matrix((vals if uv else np.ones_like(vals), A.indices, A.indptr), shape=(n, n))
        acc = M @ Z
        cab = np.array([epi_coef(epi, 0.1, x) for x in d])
        ref = cab[:, :1] * acc + cab[:, 1:] * T
        assert np.abs(out - ref).max() < 1e-9


@pytest.mark.parametrize("kw", [
    dict(n_ctas=6, warps_per_cta=4, slot_rows=40, min_hub_degree=8, fine_cols=64, coarse_edges=512),
    dict(n_ctas=3, warps_per_cta=16, slot_rows=100, min_hub_degree=4, fine_cols=32, coarse_edges=256, part_div=8),
    dict(n_ctas=1, warps_per_cta=1, slot_rows=30, min_hub_degree=16, fine_cols=128, coarse_edges=4096),
    dict(n_ctas=16, warps_per_cta=2, slot_rows=48, min_hub_degree=2, fine_cols=64, coarse_edges=64, slack=0),
])
def test_tiled_plan_walks_to_one_step_rmat(kw):
    ip, idx = oracle.rmat_graph(6000, 90000, 13, seed=1)
    A, d, vals = _ahat(ip, idx)
    tp = build_tiled_plan(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int32)),
                          torch.from_numpy(vals), rest_chunk_edges=128, **kw)
    assert tp.stats["hub_rows"] > 0 and tp.stats["slots_per_cta_max"] <= kw["slot_rows"]
    _check(A, d, vals, tp)


def test_tiled_plan_all_rows_are_hubs_on_cora():
    z, adj = load_std("cora_ml")
    A, d, vals = _ahat(z["adj_indptr"].astype(np.int64), z["adj_indices"])
    tp = build_tiled_plan(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int32)),
                          torch.from_numpy(vals), n_ctas=8, warps_per_cta=8, slot_rows=400, min_hub_degree=1, fine_cols=64,
                          coarse_edges=256, rest_chunk_edges=128)
    assert tp.rest is None and tp.stats["hub_rows"] == A.shape[0]
    _check(A, d, vals, tp)


def test_tiled_plan_value_free_without_vals():
    ip, idx = oracle.rmat_graph(3000, 40000, 12, seed=2)
    A, d, vals = _ahat(ip, idx)
    tp = build_tiled_plan(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int32)), None,
                          n_ctas=4, warps_per_cta=4, slot_rows=32, min_hub_degree=8, fine_cols=64, coarse_edges=512,
                          rest_chunk_edges=128)
    assert tp.vals is None and tp.rest.vals is None
    n = A.shape[0]
    rng = np.random.RandomState(3)
    Z, T = rng.randn(n, 3), rng.randn(n, 3)
    out = walk_tiled(tp, Z, T, 0.2, 2, False)
    rest = walk_stream(tp.rest, Z, T, 0.2, 2, False, rows=tp.rest.order.numpy())
    out = np.where(np.isnan(out), rest, out)
    M = sp.csr_matrix((np.ones_like(vals), A.indices, A.indptr), shape=(n, n))
    cab = np.array([epi_coef(2, 0.2, x) for x in d])
    assert np.abs(out - (cab[:, :1] * (M @ Z) + cab[:, 1:] * T)).max() < 1e-9


def test_choose_windows_covers_the_column_space():
    hist = torch.tensor([1000, 800, 600, 50, 40, 30, 20, 10, 5, 5, 5, 1], dtype=torch.int64)
    ends, n_fine = choose_windows(hist, bucket=32, n_ctas=2, fine_cols=64, fine_min_reuse=1.5, coarse_edges=40, n_cols=370)
    e = ends.tolist()
    assert e == sorted(set(e)) and e[-1] == 370 and e[0] > 0
    assert n_fine >= 1 and e[:n_fine] == [64 * (i + 1) for i in range(n_fine)]


def test_tiled_plan_rejects_bad_arguments():
    ip, idx = oracle.rmat_graph(500, 3000, 9, seed=2)
    A, d, vals = _ahat(ip, idx)
    a = (torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int32)))
    with pytest.raises(ValueError):
        build_tiled_plan(*a, n_ctas=2, warps_per_cta=17)
    with pytest.raises(ValueError):
        build_tiled_plan(*a, n_ctas=2, min_hub_degree=10 ** 6)
