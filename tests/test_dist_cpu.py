"""Host-side logic of the row-partitioned multi-GPU propagation (ppnp_b200/dist.py) on CPU with the
gloo backend, world sizes 2 and 3: partition by non-zero prefix, column remap, halo lists, the
all-to-all exchange and the step orchestration (interior/boundary split, Y-space epilogues).  The
local arithmetic is the numpy walker of the edge stream (tests/util.py); the CUDA kernel itself is
covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT, oracle, relerr, epi_coef


def walker_step(plan, Zin, T, Zout, alpha, epi, use_vals):
    """numpy stand-in for ppnp_spmm_step over a (subset) plan: writes only the rows the plan produces."""
    cols = plan.cols.numpy(); vals = plan.vals.numpy() if plan.vals is not None else None
    seg_row = plan.seg_row.numpy(); chunk_seg = plan.chunk_seg.numpy(); W = plan.chunk_edges
    Zi = Zin.numpy().astype(np.float64); Tn = T.numpy().astype(np.float64).copy()   # T may be Zout (EPI_ACC)
    rd = plan.row_deg.numpy() if plan.row_deg is not None else None
    acc_mode = bool(epi & 16)
    epi &= 15
    F = Zi.shape[1]
    partial = np.zeros((max(plan.n_slots, 1), F))
    for c in range(plan.n_chunks):
        s = int(chunk_seg[c]); acc = np.zeros(F); cnt = 0
        for e in range(c * W, (c + 1) * W):
            raw = int(cols[e]); col = raw & 0x7FFFFFFF
            acc += (vals[e] if use_vals else 1.0) * Zi[col]; cnt += 1
            if raw < 0:
                sv = int(seg_row[s]); s += 1
                if sv < 0:
                    partial[sv & 0x7FFFFFFF] = acc
                else:
                    a, b = epi_coef(epi, alpha, cnt if rd is None else float(rd[sv]))
                    b = 1.0 if acc_mode else b
                    Zout[sv] = torch.from_numpy((a * acc + b * Tn[sv]).astype(np.float32))
                acc = np.zeros(F); cnt = 0
    fp, fr, fd = plan.fix_ptr.numpy(), plan.fix_row.numpy(), plan.fix_deg.numpy()
    for q in range(plan.n_fix):
        a, b = epi_coef(epi, alpha, float(fd[q]))
        b = 1.0 if acc_mode else b
        Zout[fr[q]] = torch.from_numpy((a * partial[fp[q]:fp[q + 1]].sum(0) + b * Tn[fr[q]]).astype(np.float32))


def worker(rank, world, port, phases, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ppnp_b200 import dist as pd
        n, F, K, alpha = 1500, 5, 4, 0.1
        ip, idx = oracle.rmat_graph(n, 20000, 11, seed=5)
        oip, oidx, oval, odeg = oracle.c_a_hat(ip, idx, None, "sym")          # A + I structure, global columns
        deg = torch.from_numpy(np.diff(oip))
        bounds = pd.balanced_row_blocks(deg, world)
        assert bounds[0] == 0 and bounds[-1] == n and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))
        nnz_blocks = [int(oip[bounds[r + 1]] - oip[bounds[r]]) for r in range(world)]
        assert max(nnz_blocks) < 1.5 * (oip[-1] / world) + deg.max().item()   # balanced by non-zeros
        lo, hi = bounds[rank], bounds[rank + 1]
        ipl = torch.from_numpy(oip[lo:hi + 1] - oip[lo])
        colsg = torch.from_numpy(oidx[oip[lo]:oip[hi]].astype(np.int64))
        topo = pd.build_shard_topology(ipl, colsg, bounds, rank)
        # remap is invertible and the halo is exactly the set of remote columns
        ext_ids = torch.cat([torch.arange(lo, hi), topo.halo_cols])
        assert torch.equal(ext_ids[topo.indices.long()], colsg)
        assert sum(topo.recv_counts) == topo.n_halo
        dinv = pd.global_dinv(ipl, bounds, rank, world, torch.device("cpu"))
        assert np.allclose(dinv.numpy(), 1 / np.sqrt(odeg))
        if phases == "fused":
            prop = pd.FusedPushPropagation(topo, dinv, chunk_edges=128, step_fn=walker_step)
        elif phases == "fused-carve":      # hot column blocks of the shard's [local | halo] column space first
            prop = pd.FusedPushPropagation(topo, dinv, chunk_edges=128, step_fn=walker_step,
                                           carve=dict(block_cols=100, n_blocks=6, min_piece=3))
            assert prop.sub.plan.carve["carved_edges"] > 0 and not prop.sub.plan.wide_cta
        elif phases == "fused-window":     # rank-sorted rows, whole-segment chunks in column-window order
            prop = pd.FusedPushPropagation(topo, dinv, chunk_edges=128, step_fn=walker_step, window="mid")
            assert "window/mid" in prop.transport_name
        elif phases.startswith("hybrid"):  # hub rows summed where their columns live (partial rows + combine launch)
            prop = pd.HybridPushPropagation(topo, dinv, hub_degree=int(phases[6:]), chunk_edges=128, step_fn=walker_step,
                                            alpha=alpha)
            st = prop.stats
            assert st["n_halo"] <= st["n_halo_1d"]
            tot = torch.tensor([st["n_virtual"], st["n_pslots"], st["combined_rows"], st["n_halo_1d"] - st["n_halo"]])
            dist.all_reduce(tot)
            assert int(tot[0]) == int(tot[1])                              # every virtual row has exactly one slot somewhere
            if int(phases[6:]) <= 64:
                assert int(tot[0]) > 0 and int(tot[2]) > 0
            else:                                                           # no row is a hub: the plain 1-D form
                assert int(tot[0]) == 0 and int(tot[2]) == 0 and int(tot[3]) == 0
            if int(phases[6:]) <= 16:
                assert int(tot[3]) > 0                                      # the halo really shrinks
        elif phases.startswith("pipe"):
            prop = pd.PipelinedPushPropagation(topo, dinv, chunk_edges=128, row_groups=int(phases[4:]), step_fn=walker_step)
        else:
            prop = pd.PartitionedPropagation(topo, dinv, chunk_edges=128, phases=phases, step_fn=walker_step)
        rng = np.random.RandomState(0)
        Hg = rng.randn(n, F).astype(np.float32)                                # the same global H on every rank
        n_ext = prop.rows_alloc if (phases.startswith("pipe") or phases.startswith("fused") or phases.startswith("hybrid")) else prop.n_ext()
        H = torch.zeros(n_ext, F); H[: topo.n_local] = torch.from_numpy(Hg[lo:hi])
        Z, S = torch.zeros_like(H), torch.zeros_like(H)
        out = prop.propagate(H, Z, S, K, alpha)
        import scipy.sparse as sp
        A = sp.csr_matrix((oval, oidx, oip), shape=(n, n))
        ref = oracle.appnp(A, Hg.astype(np.float64), alpha, K)[lo:hi]
        err = relerr(out.numpy(), ref)
        assert err < 1e-5, err
        if phases.startswith("hybrid"):
            # a second call on the same buffers (stale partial rows from the first one must not matter), K = 2
            out2 = prop.propagate(H, Z, S, 2, alpha)
            assert relerr(out2.numpy(), oracle.appnp(A, Hg.astype(np.float64), alpha, 2)[lo:hi]) < 1e-5
            with pytest.raises(NotImplementedError):
                prop.propagate(H, Z, S, 1, alpha)
        else:
            # K = 1 path (plain epilogue)
            out1 = prop.propagate(H, Z, S, 1, alpha)
            assert relerr(out1.numpy(), oracle.appnp(A, Hg.astype(np.float64), alpha, 1)[lo:hi]) < 1e-5
        with open(os.path.join(outdir, f"ok_{rank}"), "w") as f:
            f.write(f"{err}\n")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,phases", [(2, "peer"), (2, "one"), (3, "peer"), (3, "two"), (2, "pipe3"), (3, "pipe4"), (3, "pipe1"), (2, "fused"), (3, "fused"),
                                          (2, "fused-carve"), (3, "fused-carve"), (2, "fused-window"), (3, "fused-window"),
                                          (2, "hybrid8"), (3, "hybrid40"), (2, "hybrid100000")])
def test_partitioned_propagation_gloo(tmp_path, world, phases):
    port = 29600 + world * 10 + len(phases) + (os.getpid() % 50)
    mp.spawn(worker, args=(world, port, phases, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok_{r}") for r in range(world))


def test_balanced_row_blocks_handles_skew():
    from ppnp_b200.dist import balanced_row_blocks
    w = torch.ones(1000, dtype=torch.int64)
    w[:10] = 1000
    b = balanced_row_blocks(w, 4)
    sums = [int(w[b[i]:b[i + 1]].sum()) for i in range(4)]
    assert b[0] == 0 and b[-1] == 1000 and max(sums) <= 2 * (int(w.sum()) // 4)
    assert balanced_row_blocks(torch.ones(5), 8)[-1] == 5      # more ranks than rows: empty blocks allowed


def test_stripe_relabel_is_a_bijection_that_mixes_the_id_space():
    from ppnp_b200.dist import auto_stripes, stripe_relabel
    for n, world in ((204_800, 2), (1_000_000, 8), (100_000, 4)):
        st = auto_stripes(n, world)
        assert st >= 1 and (st == 1 or n % (world * st) == 0)
        f = stripe_relabel(torch.arange(n), n, world, st)
        assert torch.equal(torch.sort(f).values, torch.arange(n))
        if st > 1:
            # the first 1/world of the NEW ids draws evenly from the whole old range
            old_in_first = torch.nonzero(f < n // world).flatten()
            assert old_in_first.max() > 0.9 * n and abs(float(old_in_first.float().mean()) / n - 0.5) < 0.1
    assert auto_stripes(1000, 1) == 1 and auto_stripes(1001, 2) == 1      # nothing to deal / not divisible
    assert torch.equal(stripe_relabel(torch.arange(10), 10, 1, 4), torch.arange(10))


def test_degree_sort_relabel_keeps_blocks_and_orders_them_by_weight():
    from ppnp_b200.dist import degree_sort_relabel
    g = torch.Generator().manual_seed(0)
    w = torch.randint(1, 50, (1000,), generator=g, dtype=torch.int32)
    bounds = [0, 130, 130, 700, 1000]                      # an empty block in the middle
    f = degree_sort_relabel(w, bounds)
    assert torch.equal(torch.sort(f).values, torch.arange(1000))           # a bijection
    old_of_new = torch.argsort(f)
    for lo, hi in zip(bounds, bounds[1:]):
        assert bool(((f[lo:hi] >= lo) & (f[lo:hi] < hi)).all())            # nobody leaves its block
        ws = w[old_of_new[lo:hi]]
        assert bool((ws[1:] <= ws[:-1]).all())                             # descending weight inside the block ...
        same = ws[1:] == ws[:-1]
        assert bool((old_of_new[lo:hi][1:][same] > old_of_new[lo:hi][:-1][same]).all())   # ... ties by id
