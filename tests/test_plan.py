"""Host-side logic of the edge-stream plan (ppnp_b200/plan.py), checked on the CPU by walking
the stream in numpy (tests/util.py walk_stream) against the oracle."""
import numpy as np
import pytest
import torch

from util import load_golden, load_std, oracle, relerr, walk_stream
from ppnp_b200.plan import build_stream_plan, degree_order


def ahat_tensors(name):
    _, adj = load_std(name)
    ah = oracle.calc_A_hat(adj, "sym")
    return ah, torch.from_numpy(ah.indptr.astype(np.int32)), torch.from_numpy(ah.indices.astype(np.int32)), \
        torch.from_numpy(ah.data.astype(np.float32))


@pytest.mark.parametrize("order", ["natural", "degree"])
@pytest.mark.parametrize("chunk", [128, 256])
def test_plan_invariants(order, chunk):
    ah, ip, idx, val = ahat_tensors("cora_ml")
    ordt = None if order == "natural" else degree_order(ip)
    p = build_stream_plan(ip, idx, val, chunk, ordt)
    cols = p.cols.numpy()
    assert p.n_chunks % 32 == 0 and len(cols) == p.n_chunks * chunk
    assert (cols[ah.nnz:] == 0).all()          # padding: column 0, no flag
    ends = np.nonzero(cols < 0)[0]
    assert len(ends) == p.n_segs
    # no segment crosses a chunk boundary; chunk_seg counts the segments before each chunk
    starts = np.concatenate([[0], ends[:-1] + 1])
    assert (starts // chunk == ends // chunk).all()
    cs = p.chunk_seg.numpy()
    for c in range(p.n_chunks):
        assert cs[c] == np.searchsorted(starts, c * chunk, side="left")
    # partial slots are numbered in stream order and grouped by row
    sr = p.seg_row.numpy()[:p.n_segs]
    slots = sr[sr < 0] & 0x7FFFFFFF
    assert np.array_equal(slots, np.arange(p.n_slots))
    fp = p.fix_ptr.numpy()
    assert fp[0] == 0 and fp[-1] == p.n_slots and (np.diff(fp) >= 2).all()
    finals = sr[sr >= 0]
    assert len(set(finals.tolist()) | set(p.fix_row.numpy().tolist())) == p.n
    # split rows are exactly the rows longer than the room left in their chunk
    deg = np.diff(ah.indptr)
    assert (p.fix_deg.numpy() == deg[p.fix_row.numpy()]).all()


@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
@pytest.mark.parametrize("order", ["natural", "degree"])
def test_stream_walk_matches_oracle_one_step(name, order):
    ah, ip, idx, val = ahat_tensors(name)
    g = load_golden(name)
    H = g["H"].astype(np.float64)
    ordt = None if order == "natural" else degree_order(ip)
    p = build_stream_plan(ip, idx, val, 128, ordt)
    ref = oracle.appnp(ah, H, 0.1, 1)
    got = walk_stream(p, H, H, 0.1, epi=0, use_vals=True)
    assert relerr(got, ref) < 1e-7   # fp32 stored values


def test_value_free_y_space_equals_stored_values():
    """K steps: Z2Y (stored values) -> Y ... -> Y2Z (value-free) must equal the plain iteration."""
    ah, ip, idx, val = ahat_tensors("citeseer")
    g = load_golden("citeseer")
    H = g["H"].astype(np.float64)
    p = build_stream_plan(ip, idx, val, 128, None)
    K, alpha = 4, 0.1
    Z = walk_stream(p, H, H, alpha, epi=1, use_vals=True)
    for k in range(2, K):
        Z = walk_stream(p, Z, H, alpha, epi=2, use_vals=False)
    Z = walk_stream(p, Z, H, alpha, epi=3, use_vals=False)
    assert relerr(Z, oracle.appnp(ah, H, alpha, K)) < 1e-7


def test_rw_mode_value_free():
    _, adj = load_std("citeseer")
    ah = oracle.calc_A_hat(adj, "rw")
    ip = torch.from_numpy(ah.indptr.astype(np.int32))
    idx = torch.from_numpy(ah.indices.astype(np.int32))
    H = np.random.RandomState(0).randn(adj.shape[0], 3)
    p = build_stream_plan(ip, idx, None, 128, None)
    got = walk_stream(p, H, H, 0.2, epi=4, use_vals=False)
    assert relerr(got, oracle.appnp(ah, H, 0.2, 1)) < 1e-12


def test_hub_row_is_split_and_summed_in_order():
    # a star: row 0 has 1000 neighbours -> many partial segments
    import scipy.sparse as sp
    n = 1001
    rows = np.concatenate([np.zeros(1000, dtype=int), np.arange(1, n)])
    cols = np.concatenate([np.arange(1, n), np.zeros(1000, dtype=int)])
    adj = sp.csr_matrix((np.ones(2000, dtype=np.float32), (rows, cols)), shape=(n, n))
    ah = oracle.calc_A_hat(adj, "sym")
    p = build_stream_plan(torch.from_numpy(ah.indptr.astype(np.int32)), torch.from_numpy(ah.indices.astype(np.int32)),
                          torch.from_numpy(ah.data.astype(np.float32)), 128, None)
    assert p.n_fix >= 1 and 0 in p.fix_row.tolist() and p.n_slots >= 8
    H = np.random.RandomState(1).randn(n, 5)
    assert relerr(walk_stream(p, H, H, 0.1, 0, True), oracle.appnp(ah, H, 0.1, 1)) < 1e-7


def test_rejects_rows_without_self_loop():
    ip = torch.tensor([0, 1, 1], dtype=torch.int32)
    idx = torch.tensor([0], dtype=torch.int32)
    with pytest.raises(ValueError):
        build_stream_plan(ip, idx, None, 128, None)


# ------------------------------------------------------------------ carved streams / lane-transposed layout
from ppnp_b200.plan import build_carved_plan, lane_transpose, lane_group_for  # noqa: E402


@pytest.mark.parametrize("name,block_cols,n_blocks,min_piece",
                         [("citeseer", 16, 8, 2), ("citeseer", 64, 4, 3), ("cora_ml", 512, 64, 4)])
def test_carved_stream_walk_matches_oracle(name, block_cols, n_blocks, min_piece):
    ah, ip, idx, val = ahat_tensors(name)
    g = load_golden(name)
    H = g["H"].astype(np.float64)
    p = build_carved_plan(ip, idx, val, 128, block_cols, n_blocks, min_piece)
    assert p.wide_cta and p.carve["carved_edges"] > 0
    ref = oracle.appnp(ah, H, 0.1, 1)
    assert relerr(walk_stream(p, H, H, 0.1, epi=0, use_vals=True), ref) < 1e-7
    # value-free Y-space steps use the row degree: fix_deg for split rows, segment length otherwise
    K, alpha = 3, 0.1
    Z = walk_stream(p, H, H, alpha, epi=1, use_vals=True)
    Z = walk_stream(p, Z, H, alpha, epi=2, use_vals=False)
    Z = walk_stream(p, Z, H, alpha, epi=3, use_vals=False)
    assert relerr(Z, oracle.appnp(ah, H, alpha, K)) < 1e-7


def test_carved_stream_structure():
    ah, ip, idx, val = ahat_tensors("cora_ml")
    W, BC, NB, T = 128, 32, 6, 3
    p = build_carved_plan(ip, idx, val, W, BC, NB, T)
    cols = p.cols.numpy()
    ends = np.nonzero(cols < 0)[0]
    starts = np.concatenate([[0], ends[:-1] + 1])
    assert len(ends) == p.n_segs and (starts // W == ends // W).all()
    assert np.array_equal(p.chunk_seg.numpy(), np.searchsorted(starts, np.arange(p.n_chunks) * W, side="left"))
    # every edge of A_hat appears exactly once
    deg = np.diff(ah.indptr)
    sr = p.seg_row.numpy()[:p.n_segs]
    fp, fr = p.fix_ptr.numpy(), p.fix_row.numpy()
    slot_row = np.repeat(fr, np.diff(fp))
    seg_rows = np.where(sr < 0, slot_row[np.where(sr < 0, sr & 0x7FFFFFFF, 0)], sr)
    rows_of_edges = np.repeat(seg_rows, ends - starts + 1)
    c = cols[:ah.nnz] & 0x7FFFFFFF
    got = np.lexsort((c, rows_of_edges))
    want_rows = np.repeat(np.arange(p.n), deg)
    assert np.array_equal(rows_of_edges[got], want_rows) and np.array_equal(c[got], ah.indices)
    # slots: each used once, contiguous per fix row; fix rows have >= 2 segments, others exactly one final
    slots = np.sort(sr[sr < 0] & 0x7FFFFFFF)
    assert np.array_equal(slots, np.arange(p.n_slots)) and (np.diff(fp) >= 2).all()
    finals = sr[sr >= 0]
    assert len(finals) + p.n_fix == p.n and len(set(finals.tolist()) & set(fr.tolist())) == 0
    assert (p.fix_deg.numpy() == deg[fr]).all()
    # the carved part comes first, block by block: column ranks of carved pieces stay inside their block
    order = degree_order(ip).numpy()
    rank = np.empty(p.n, dtype=np.int64)
    rank[order] = np.arange(p.n)
    n_carved = p.carve["carved_edges"]
    blk = rank[c[:n_carved]] // BC
    assert (np.diff(blk) >= 0).all() and blk.max() < NB
    # pieces honour min_piece before chunk cuts: count edges per (row, block) in the carved part
    key = rows_of_edges[:n_carved] * NB + blk
    _, cnt = np.unique(key, return_counts=True)
    assert cnt.min() >= T


def test_carve_nothing_equals_degree_order_stream():
    ah, ip, idx, val = ahat_tensors("citeseer")
    a = build_carved_plan(ip, idx, val, 128, 64, 0, 4)
    b = build_stream_plan(ip, idx, val, 128, degree_order(ip))
    assert torch.equal(a.cols, b.cols) and torch.equal(a.vals, b.vals) and torch.equal(a.chunk_seg, b.chunk_seg)
    assert a.n_slots == b.n_slots and a.n_fix == b.n_fix


@pytest.mark.parametrize("G", [4, 8, 16, 32])
def test_lane_transposed_layout_walks_to_the_same_result(G):
    ah, ip, idx, val = ahat_tensors("citeseer")
    H = np.random.RandomState(2).randn(ah.shape[0], 3)
    p = build_stream_plan(ip, idx, val, 128, degree_order(ip))
    q = lane_transpose(p, G)
    assert q.lane_group == G and (q.struct().flags >> 8) == G
    assert not torch.equal(q.cols, p.cols)
    # the 4 words a lane stages with one 16-byte copy are the words it consumes over 4/SR slabs
    SR = max(1, 16 // G); SE = SR * G; CPS = 4 // SR
    c0, q0 = p.cols.numpy()[:128], q.cols.numpy()[:128]
    for quad in range(128 // (CPS * SE)):
        for lane in range(G):
            mine = [c0[(quad * CPS + s) * SE + r * G + lane] for s in range(CPS) for r in range(SR)]
            base = quad * CPS * SE + lane * 4
            assert list(q0[base:base + 4]) == mine
    a = walk_stream(p, H, H, 0.1, 0, True)
    b = walk_stream(q, H, H, 0.1, 0, True)
    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        lane_transpose(q, G)


def test_lane_group_for_matches_dispatch():
    assert [lane_group_for(F) for F in (64, 16, 32, 128, 256, 7, 3, 1, 48)] == [16, 4, 8, 32, 32, 8, 4, 1, 16]


def test_interleaved_carved_stream_is_a_chunk_permutation_with_the_same_result():
    ah, ip, idx, val = ahat_tensors("cora_ml")
    H = np.random.RandomState(5).randn(ah.shape[0], 2)
    a = build_carved_plan(ip, idx, val, 128, 64, 8, 3)
    b = build_carved_plan(ip, idx, val, 128, 64, 8, 3, interleave=True, unit_chunks=4)
    assert b.carve["interleave"] and not torch.equal(a.cols, b.cols)
    ca, cb = a.cols.view(a.n_chunks, 128), b.cols.view(b.n_chunks, 128)
    # same multiset of chunks, each still pointing at its own segments
    ka = sorted(zip(a.chunk_seg.tolist(), [tuple(r) for r in ca.tolist()]))
    kb = sorted(zip(b.chunk_seg.tolist(), [tuple(r) for r in cb.tolist()]))
    assert ka == kb
    # carved and residual units alternate: the first residual chunk comes long before the last carved one
    lead = a.carve["carved_edges"] // 128
    where = {tuple(r): i for i, r in enumerate(cb.tolist())}
    real = (a.nnz + 127) // 128                      # chunks after this one are identical padding
    pos = np.array([where[tuple(r)] for r in ca.tolist()[:real]])
    assert pos[lead:].min() < pos[:lead].max() and (np.diff(pos[:lead]) > 0).all() and (np.diff(pos[lead:]) > 0).all()
    za = walk_stream(a, H, H, 0.1, 0, True)
    zb = walk_stream(b, H, H, 0.1, 0, True)
    assert np.array_equal(za, zb)


def test_two_level_carve_blocks_stack_along_the_rank_axis():
    ah, ip, idx, val = ahat_tensors("cora_ml")
    H = np.random.RandomState(6).randn(ah.shape[0], 2)
    one = build_carved_plan(ip, idx, val, 128, 32, 4, 3)
    same = build_carved_plan(ip, idx, val, 128, levels=[(32, 4, 3)])
    assert torch.equal(one.cols, same.cols) and torch.equal(one.seg_row, same.seg_row)
    two = build_carved_plan(ip, idx, val, 128, levels=[(32, 4, 3), (500, 3, 6)])
    assert two.carve["n_blocks"] == 7 and two.carve["carved_edges"] > one.carve["carved_edges"]
    # carved part: block index never decreases; blocks 0-3 are 32 ranks wide, 4-6 are 500 wide
    order = degree_order(ip).numpy()
    rank = np.empty(two.n, dtype=np.int64); rank[order] = np.arange(two.n)
    r = rank[two.cols.numpy()[:two.carve["carved_edges"]] & 0x7FFFFFFF]
    blk = np.where(r < 128, r // 32, 4 + (r - 128) // 500)
    assert (np.diff(blk) >= 0).all() and blk.max() == 6
    assert relerr(walk_stream(two, H, H, 0.1, 0, True), oracle.appnp(ah, H, 0.1, 1)) < 1e-7


def test_window_order_is_a_chunk_permutation_of_rank_sorted_rows_with_the_same_result():
    """order="window": rows in degree order, a row's columns hottest first, whole-segment chunks by column window.
    Cora-ML has one hub (246 neighbours); a synthetic skewed graph supplies many whole-segment chunks."""
    from ppnp_b200.plan import rank_sorted_csr, window_order_chunks, column_ranks
    for name, W in (("cora_ml", 128), ("rmat", 128)):
        if name == "rmat":
            rip, ridx = oracle.rmat_graph(3000, 200000, 12, seed=3)
            oip, oidx, oval, _ = oracle.c_a_hat(rip, ridx, None, "sym")
            import scipy.sparse as sp
            ah = sp.csr_matrix((oval, oidx, oip), shape=(len(oip) - 1,) * 2)
            ip, idx, val = torch.from_numpy(oip.astype(np.int32)), torch.from_numpy(oidx.astype(np.int32)), torch.from_numpy(oval.astype(np.float32))
        else:
            ah, ip, idx, val = ahat_tensors(name)
        n = ah.shape[0]
        H = np.random.RandomState(8).randn(n, 3)
        sidx, sval, crank = rank_sorted_csr(ip, idx, val)
        assert torch.equal(crank, column_ranks(ip, idx))
        # same triples per row, columns by ascending rank
        ipl = ip.to(torch.int64)
        for r in (0, 1, int(torch.argmax(ipl[1:] - ipl[:-1]))):
            a, b = int(ipl[r]), int(ipl[r + 1])
            assert sorted(idx[a:b].tolist()) == sorted(sidx[a:b].tolist())
            assert (np.diff(crank[sidx[a:b].to(torch.int64)].numpy()) > 0).all()
        base = build_stream_plan(ip, sidx, sval, W, degree_order(ip))
        win = window_order_chunks(base, crank, key="mid")
        cb, cw = base.cols.view(base.n_chunks, W), win.cols.view(win.n_chunks, W)
        assert sorted(zip(base.chunk_seg.tolist(), [tuple(r) for r in cb.tolist()])) == sorted(zip(win.chunk_seg.tolist(), [tuple(r) for r in cw.tolist()]))
        single = ((cw[:, :-1] >= 0).all(1) & (cw[:, -1] < 0)).numpy()
        k = int(single.sum())
        if name == "rmat":
            assert k > 50 and not torch.equal(base.cols, win.cols)
        assert single[:k].all()                                   # they lead the stream ...
        mid = crank[(cw[:k, W // 2] & 0x7FFFFFFF).to(torch.int64)].numpy()
        assert (np.diff(mid) >= 0).all()                          # ... sorted by the rank of their middle column
        ref = oracle.appnp(ah, H, 0.1, 1)
        for use_vals, epi in ((True, 0),):
            assert relerr(walk_stream(win, H, H, 0.1, epi, use_vals), ref) < 1e-7
        assert np.array_equal(walk_stream(win, H, H, 0.1, 0, True), walk_stream(base, H, H, 0.1, 0, True))
        # lane transposition composes with it
        wt = lane_transpose(win, 16)
        assert relerr(walk_stream(wt, H, H, 0.1, 0, True), ref) < 1e-7


def test_rank_sorted_csr_batches_give_the_same_arrays():
    """The row-block batching of rank_sorted_csr (bounded temporaries next to a multi-GB shard) must not show in the result."""
    from ppnp_b200.plan import rank_sorted_csr
    ah, ip, idx, val = ahat_tensors("citeseer")
    full = rank_sorted_csr(ip, idx, val)
    for mb in (1, 7, 100, 5000):            # 1: one row per batch (rows longer than the batch still go whole)
        part = rank_sorted_csr(ip, idx, val, max_batch_edges=mb)
        assert torch.equal(part[0], full[0]) and torch.equal(part[1], full[1]) and torch.equal(part[2], full[2])
    # rectangular column space (a shard: local rows x [local | halo] columns): ranks by reference count
    n = ah.shape[0]
    wide = rank_sorted_csr(ip, idx, None, n_cols=n + 50)
    assert wide[2].numel() == n + 50 and torch.equal(torch.sort(wide[2]).values, torch.arange(n + 50))
    refs = torch.bincount(idx.to(torch.int64), minlength=n + 50)
    assert bool((refs[torch.argsort(wide[2])][1:] <= refs[torch.argsort(wide[2])][:-1]).all())
