"""Property tests of the edge-stream builders (ppnp_b200/plan.py) on random small graphs: whatever the
processing order, carving, interleaving, chunk size or lane layout, walking the stream (tests/util.py,
a literal numpy reading of what csrc/appnp_spmm.cu does with the arrays) gives (A + I) Y row by row."""
import numpy as np
import scipy.sparse as sp
import torch
from hypothesis import given, settings, strategies as st

from util import walk_stream
from ppnp_b200.plan import build_carved_plan, build_stream_plan, degree_order, interleave_chunks, lane_transpose


def random_pattern(seed, n, density, hub):
    rng = np.random.RandomState(seed)
    a = rng.rand(n, n) < density
    if hub:
        a[0, :] = True          # one row that spans several chunks
    a = a | a.T
    np.fill_diagonal(a, True)   # A + I: every row holds its self loop
    m = sp.csr_matrix(a.astype(np.float32))
    m.sort_indices()
    return m


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 70), density=st.floats(0.0, 0.4), hub=st.booleans(),
       chunk=st.sampled_from([128, 256]), block_cols=st.integers(1, 40), n_blocks=st.integers(0, 6),
       min_piece=st.integers(1, 6), interleave=st.booleans(), unit=st.integers(1, 3), lane_group=st.sampled_from([0, 4, 8, 16, 32]),
       two_level=st.booleans())
def test_any_stream_computes_the_same_product(seed, n, density, hub, chunk, block_cols, n_blocks, min_piece, interleave, unit,
                                              lane_group, two_level):
    m = random_pattern(seed, n, density, hub)
    ip, idx = torch.from_numpy(m.indptr.astype(np.int32)), torch.from_numpy(m.indices.astype(np.int32))
    val = torch.from_numpy(np.random.RandomState(seed + 1).rand(m.nnz).astype(np.float32))
    Y = np.random.RandomState(seed + 2).randn(n, 3)
    T = np.random.RandomState(seed + 3).randn(n, 3)
    mv = sp.csr_matrix((val.numpy().astype(np.float64), m.indices, m.indptr), shape=m.shape)
    want_vals = 0.9 * (mv @ Y) + 0.1 * T                                        # PPNP_EPI_PLAIN, stored values
    deg = np.diff(m.indptr)[:, None].astype(np.float64)
    want_free = 0.9 / deg * (m.astype(np.float64) @ Y) + 0.1 / np.sqrt(deg) * T   # PPNP_EPI_Y, value-free
    levels = [(block_cols, n_blocks, min_piece), (block_cols * 3, 2, min_piece + 1)] if two_level else None
    plans = [build_stream_plan(ip, idx, val, chunk, None), build_stream_plan(ip, idx, val, chunk, degree_order(ip)),
             build_carved_plan(ip, idx, val, chunk, block_cols, n_blocks, min_piece, interleave=interleave, unit_chunks=unit,
                               levels=levels)]
    plans.append(interleave_chunks(plans[1], plans[1].n_chunks // 2, unit))
    if lane_group:
        plans = [lane_transpose(p, lane_group) for p in plans]
    for p in plans:
        assert p.n_chunks % 32 == 0 and p.cols.numel() == p.n_chunks * chunk
        assert np.allclose(walk_stream(p, Y, T, 0.1, 0, True), want_vals, rtol=1e-6, atol=1e-6)
        assert np.allclose(walk_stream(p, Y, T, 0.1, 2, False), want_free, rtol=1e-9, atol=1e-9)
