"""BASELINE configs 2 and 3 at their own size: a PubMed-shape graph (n = 19 717, ~88.6 k stored entries of A; the data
file is absent from the reference, .MISSING_LARGE_BLOBS, so the graph is the seeded power-law synthetic bench.py
uses).  Pi against the fp64 oracle on probed rows (helpers.py:68-71 restated; a dense fp64 inverse of this size does
not finish in seconds, rows of the symmetric Pi from unit vectors do: KAT-2 of SURVEY 8c), the two gather-GEMMs at
main.py's shapes (model.py:63), the literal top-k lines and one literal batch (batch-main.py:113-117, 140-146) on the
Pi the GPU built.  Needs a GPU."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from util import oracle, relerr

pytestmark = pytest.mark.gpu
N, NNZ_A, C, ALPHA = 19717, 88648, 3, 0.1


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def world():
    import ppnp_b200 as P
    from ppnp_b200.synth import powerlaw_adjacency
    ip, idx = powerlaw_adjacency(N, NNZ_A, seed=0, device=dev())
    ahat = P.csr_normalize(ip, idx)
    adj = sp.csr_matrix((np.ones(int(ip[-1]), np.float32), idx.cpu().numpy(), ip.cpu().numpy()), shape=(N, N))
    A64 = oracle.calc_A_hat(adj, "sym")
    # bit-exact structure and degrees at this size too
    assert np.array_equal(ahat.indptr.cpu().numpy(), A64.indptr) and np.array_equal(ahat.indices.cpu().numpy(), A64.indices)
    probe = np.random.RandomState(5).choice(N, 16, replace=False)
    E = np.zeros((N, len(probe)))
    E[probe, np.arange(len(probe))] = 1.0
    rows_ref = oracle.appnp(A64, E, ALPHA, 400).T                 # (1 - alpha)^400 ~ 5e-19: rows of Pi to round-off
    Pi = P.ppr_dense(ahat, ALPHA, tol=1e-7, method="chebyshev")
    H = np.random.RandomState(1).randn(N, C).astype(np.float32)
    return dict(P=P, ahat=ahat, A64=A64, probe=probe, rows_ref=rows_ref, Pi=Pi, H=H)


def test_pi_rows_match_the_oracle_both_methods(world):
    P, probe, ref = world["P"], world["probe"], world["rows_ref"]
    t = torch.from_numpy(probe).to(dev())
    got = world["Pi"][t].cpu().numpy().astype(np.float64)
    assert relerr(got, ref) < 1e-5
    assert np.abs(got - ref).max() < 2e-6                          # entries are O(0.1) on the diagonal, 1e-9 far away
    Pw = P.ppr_dense(world["ahat"], ALPHA, tol=1e-7, method="power")
    assert relerr(Pw[t].cpu().numpy().astype(np.float64), ref) < 1e-5
    assert float((Pw - world["Pi"]).norm() / Pw.norm()) < 1e-6
    # symmetric and strictly positive on the probed rows' diagonal
    assert float((world["Pi"][t][:, t] - world["Pi"][t][:, t].T).abs().max()) < 1e-7
    del Pw


@pytest.mark.parametrize("m", [60, 500, 940])
def test_forward_and_adjoint_at_main_py_shapes(world, m):
    """model.py:63 ``ppr[idx] @ H`` for the train / stopping / validation index sets of PubMed (60 / 500 / ~940 rows)."""
    P, Pi = world["P"], world["Pi"]
    idx = np.sort(np.random.RandomState(m).choice(N, m, replace=False))
    # force the probed rows into the index set: their oracle rows give an independent answer
    idx[: len(world["probe"])] = world["probe"]
    idx = np.unique(idx)
    ti = torch.from_numpy(idx).to(dev())
    H = torch.from_numpy(world["H"]).to(dev()).requires_grad_(True)
    out = P.ppr_matmul(Pi, H, ti)
    G64 = Pi[ti].cpu().numpy().astype(np.float64)                  # the gathered rows, as the reference materialises them
    ref = G64 @ world["H"].astype(np.float64)
    assert relerr(out.detach().cpu().numpy(), ref) < 1e-5
    assert (out.detach().cpu().numpy().argmax(1) == ref.argmax(1)).all()
    where = {r: i for i, r in enumerate(idx.tolist())}
    pr = [where[r] for r in world["probe"].tolist()]
    assert relerr(out.detach().cpu().numpy()[pr], world["rows_ref"] @ world["H"].astype(np.float64)) < 1e-5
    Gn = np.random.RandomState(m + 1).randn(len(idx), C).astype(np.float32)
    out.backward(torch.from_numpy(Gn).to(dev()))
    assert relerr(H.grad.cpu().numpy(), G64.T @ Gn.astype(np.float64)) < 1e-5
    # tcgen05 bf16 operands, fp32 accumulate: 1e-2 (north_star)
    ob = P.gather_gemm_bf16(P.to_bf16_padded(Pi), H.detach(), ti)
    assert relerr(ob.cpu().numpy(), ref) < 1e-2


def test_full_matrix_apply_bf16_and_fp32(world):
    """Config 2's benchmark shape: every row of Pi, N = 3 and 64 right-hand sides."""
    P, Pi = world["P"], world["Pi"]
    Pb = P.to_bf16_padded(Pi)
    for Ncols in (3, 64):
        Hn = np.random.RandomState(Ncols).randn(N, Ncols).astype(np.float32)
        H = torch.from_numpy(Hn).to(dev())
        ref = world["rows_ref"] @ Hn.astype(np.float64)
        t = torch.from_numpy(world["probe"]).to(dev())
        assert relerr(P.gather_gemm(Pi, H, None)[t].cpu().numpy(), ref) < 1e-5
        assert relerr(P.gather_gemm_bf16(Pb, H, None)[t].cpu().numpy(), ref) < 1e-2
        full32 = P.gather_gemm(Pi, H, None)
        assert float((P.gather_gemm_bf16(Pb, H, None) - full32).norm() / full32.norm()) < 1e-2


@pytest.mark.parametrize("k", [128, 256])
def test_topk_and_one_literal_batch(world, k):
    P = world["P"]
    Pi = world["Pi"].clone()
    th = P.topk_thresh(Pi, k).cpu().numpy()
    # literal batch-main.py:115 on sampled rows (numpy partition: exact selection, like torch.topk's values)
    rows = np.random.RandomState(k).choice(N, 256, replace=False)
    host_rows = Pi[torch.from_numpy(rows).to(dev())].cpu().numpy()
    assert np.array_equal(th[rows], oracle.topk_thresh(host_rows, k))
    P.topk_sparsify_(Pi, k)
    spp = P.dense_to_sparse_ppr(Pi)
    # batch-main.py:116's broadcast: entry (i, j) survives iff ppr[i, j] >= thresh[j]
    orig = world["Pi"][torch.from_numpy(rows).to(dev())].cpu().numpy()
    kept = Pi[torch.from_numpy(rows).to(dev())].cpu().numpy()
    assert np.array_equal(kept, np.where(orig < th[None, :], 0.0, orig).astype(np.float32))
    ipc = spp.indptr.cpu().numpy()
    assert np.array_equal((ipc[1:] - ipc[:-1])[rows], (kept > 0).sum(1))
    for B, seed in ((32, 0), (128, 1), (1024, 2)):
        idx_b = np.sort(np.random.RandomState(seed).choice(N, B, replace=False))
        ib = torch.from_numpy(idx_b).to(dev())
        dense_rows = Pi[ib].cpu().numpy()
        # oracle.batch_step indexes a dense matrix by idx_batch: hand it the B gathered rows with idx = arange(B)
        logits_ref, sel_ref = oracle.batch_step(dense_rows, np.arange(B), world["H"].astype(np.float64))
        sel = P.batch_support(spp, ib)
        assert np.array_equal(sel.cpu().numpy(), sel_ref)
        out = P.batch_propagate(spp, ib, sel, torch.from_numpy(world["H"]).to(dev())[sel])
        assert relerr(out.cpu().numpy(), logits_ref) < 1e-5
        assert (out.cpu().numpy().argmax(1) == logits_ref.argmax(1)).all()
