"""The CUDA plan builder (csrc/plan_build.cu: ppnp_plan_measure / ppnp_plan_fill) against its specification, the
tensor-op builder of ppnp_b200/plan.py, array by array and bit for bit; then through the propagation against the
C oracle.  Needs a GPU."""
import numpy as np
import pytest
import torch

from util import load_std, oracle, relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def same_plan(a, b):
    assert (a.n, a.nnz, a.chunk_edges, a.n_chunks, a.n_slots, a.n_segs, a.n_fix, a.lane_group) == \
           (b.n, b.nnz, b.chunk_edges, b.n_chunks, b.n_slots, b.n_segs, b.n_fix, b.lane_group)
    for f in ("cols", "seg_row", "chunk_seg", "fix_ptr", "fix_row", "fix_deg"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert (a.vals is None) == (b.vals is None) and (a.vals is None or torch.equal(a.vals, b.vals))
    assert (a.order is None) == (b.order is None) and (a.order is None or torch.equal(a.order, b.order))


def rmat_csr(n, raw, scale, seed):
    ip, idx = oracle.rmat_graph(n, raw, scale, seed=seed)
    oip, oidx, oval, _ = oracle.c_a_hat(ip, idx, None, "sym")
    return (torch.from_numpy(oip.astype(np.int32)).to(dev()), torch.from_numpy(oidx.astype(np.int32)).to(dev()),
            torch.from_numpy(oval.astype(np.float32)).to(dev()), (oip, oidx, oval))


@pytest.mark.parametrize("chunk", [128, 256, 512])
@pytest.mark.parametrize("order", ["natural", "degree", "subset", "shuffled"])
def test_cuda_builder_equals_tensor_op_builder(chunk, order):
    from ppnp_b200.plan import build_stream_plan_cuda, build_stream_plan_torch, degree_order, lane_transpose
    ip, idx, val, _ = rmat_csr(60000, 1500000, 16, 0)          # hub rows of thousands of edges, isolated rows, ragged tail
    n = ip.numel() - 1
    kw = {}
    if order == "degree":
        kw["order"] = degree_order(ip)
    elif order == "subset":                                    # the rows of degree >= 8, as the partitioned path lists them
        full = degree_order(ip)
        deg = (ip[1:] - ip[:-1]).to(torch.int64)
        kw = dict(order=full[: int((deg >= 8).sum())], subset=True, row_deg=deg.to(torch.float32) + 2.0)
    elif order == "shuffled":
        kw["order"] = torch.randperm(n, device=dev(), generator=torch.Generator(device=dev()).manual_seed(1))
    for v in (val, None):
        t = build_stream_plan_torch(ip, idx, v, chunk, **kw)
        c = build_stream_plan_cuda(ip, idx, v, chunk, **kw)
        same_plan(c, t)
        for G in (4, 16):
            same_plan(build_stream_plan_cuda(ip, idx, v, chunk, lane_group=G, **kw), lane_transpose(t, G))
        if order == "subset":
            assert torch.equal(c.row_deg, t.row_deg)


@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
def test_cuda_builder_on_the_reference_graphs(name):
    import ppnp_b200 as P
    from ppnp_b200.plan import build_stream_plan_cuda, build_stream_plan_torch, degree_order
    z, adj = load_std(name)
    ahat = P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev()), torch.from_numpy(z["adj_indices"]).to(dev()))
    for chunk in (128, 256):
        for o in (None, degree_order(ahat.indptr)):
            same_plan(build_stream_plan_cuda(ahat.indptr, ahat.indices, ahat.val32, chunk, o),
                      build_stream_plan_torch(ahat.indptr, ahat.indices, ahat.val32, chunk, o))


def test_cuda_builder_rejects_bad_input():
    from ppnp_b200.plan import build_stream_plan_cuda
    ip = torch.tensor([0, 2, 2, 3], dtype=torch.int32, device=dev())     # row 1 has no edge
    idx = torch.tensor([0, 1, 2], dtype=torch.int32, device=dev())
    with pytest.raises(ValueError, match="at least one edge"):
        build_stream_plan_cuda(ip, idx, None, 128)
    with pytest.raises(ValueError, match="multiple of 128"):
        build_stream_plan_cuda(ip, idx, None, 100)
    ip2 = torch.tensor([0, 2, 3, 4], dtype=torch.int32, device=dev())
    with pytest.raises(ValueError, match="indptr"):
        build_stream_plan_cuda(ip2, idx, None, 128)
    with pytest.raises(ValueError, match="order must list every row"):
        build_stream_plan_cuda(torch.tensor([0, 1, 2, 3], dtype=torch.int32, device=dev()), idx, None, 128,
                               order=torch.tensor([0, 1], device=dev()))


def test_default_path_uses_the_cuda_builder_and_matches_the_oracle(monkeypatch):
    import ppnp_b200 as P
    from ppnp_b200 import plan as plan_mod
    calls = []
    real = plan_mod.build_stream_plan_cuda
    monkeypatch.setattr(plan_mod, "build_stream_plan_cuda", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    ip, idx, val, (oip, oidx, oval) = rmat_csr(50000, 1200000, 16, 2)
    ahat = P.csr_normalize(*[torch.from_numpy(x).to(dev()) for x in (lambda g: (g[0].astype(np.int32), g[1]))(oracle.rmat_graph(50000, 1200000, 16, seed=2))])
    graph = P.PropagationGraph(ahat, order="degree", idx16=True)
    assert calls, "PropagationGraph did not go through the CUDA plan builder"
    Hn = np.random.RandomState(0).randn(50000, 64).astype(np.float32)
    ref = oracle.c_appnp_f64(oip, oidx, oval, Hn.astype(np.float64), 10, 0.1)
    Z = P.appnp_propagate(graph, torch.from_numpy(Hn).to(dev()), 10, 0.1).cpu().numpy()
    assert relerr(Z, ref) < 1e-5
