"""The bf16 tensor-core (tcgen05 / TMEM) form of the dense PPR apply against the reference's fp32
result, tolerance 1e-2 (north_star).  Needs a B200: pytest -m gpu."""
import numpy as np
import pytest
import torch

from util import load_golden, load_std, oracle, relerr

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_f32_to_bf16_rounds_to_nearest_even():
    import ppnp_b200 as P
    x = torch.randn(100003, device=dev()) * 3
    x[:4] = torch.tensor([0.0, -0.0, float("inf"), 1.0039062], device=dev())
    assert torch.equal(P.to_bf16(x), x.to(torch.bfloat16))


@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
def test_gather_gemm_bf16_matches_reference_forward(name):
    import ppnp_b200 as P
    _, adj = load_std(name)
    g = load_golden(name)
    ppr32 = torch.from_numpy(oracle.compute_ppr(adj, 0.1).astype(np.float32)).to(dev())
    Pb = P.to_bf16_padded(ppr32)
    H = torch.from_numpy(g["H"]).to(dev())
    for key, idxk in (("logits_train", "idx_train"), ("logits_full", None)):
        idx = None if idxk is None else torch.from_numpy(g[idxk]).to(dev())
        out = P.gather_gemm_bf16(Pb, H, idx).cpu().numpy()
        assert relerr(out, g[key]) < 1e-2
        # against the same bf16-rounded operands in fp64 the kernel must be exact to fp32 accumulation
        Pq = Pb.float().cpu().numpy().astype(np.float64)
        Hq = H.to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
        ref = (Pq if idx is None else Pq[g[idxk]]) @ Hq
        assert relerr(out, ref) < 1e-5


@pytest.mark.parametrize("C", [1, 7, 16, 17, 40, 64, 100])
@pytest.mark.parametrize("m,n", [(1, 64), (60, 1000), (129, 1237), (700, 2500)])
def test_gather_gemm_bf16_shapes(C, m, n):
    import ppnp_b200 as P
    rng = np.random.RandomState(C * 7 + m)
    Pi = torch.from_numpy(rng.rand(max(n, 50), n).astype(np.float32)).to(dev())
    H = torch.from_numpy(rng.randn(n, C).astype(np.float32)).to(dev())
    idx = torch.from_numpy(rng.choice(Pi.shape[0], m, replace=True)).to(dev())
    Pb = P.to_bf16_padded(Pi)
    out = P.gather_gemm_bf16(Pb, H, idx).cpu().numpy()
    Pq = Pb.float().cpu().numpy().astype(np.float64)
    Hq = H.to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
    assert relerr(out, Pq[idx.cpu().numpy()] @ Hq) < 1e-5


def test_gather_gemm_bf16_large_split_k_is_deterministic():
    import ppnp_b200 as P
    n = 6000
    Pi = torch.rand(n, n, device=dev())
    H = torch.randn(n, 7, device=dev())
    Pb = P.to_bf16_padded(Pi)
    idx = torch.randperm(n, device=dev())[:140]
    a = P.gather_gemm_bf16(Pb, H, idx)
    b = P.gather_gemm_bf16(Pb, H, idx)
    assert torch.equal(a, b)
    ref = Pb[idx].float().double() @ H.to(torch.bfloat16).double()
    assert float((a.double() - ref).norm() / ref.norm()) < 1e-5
