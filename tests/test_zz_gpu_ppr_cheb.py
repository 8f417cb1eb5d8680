"""Chebyshev-accelerated build of the dense PPR matrix (csrc/ppr_dense.cu ppnp_ppr_dense_cheb): same matrix
as helpers.py:68-71 ``compute_ppr`` in ~4x fewer steps.  Written after the round's GPU budget was spent:
these tests have not run on a GPU yet (the file sorts last so that a failure here hides nothing else)."""
import numpy as np
import pytest
import torch

from util import load_golden, load_std, relerr

NAMES = ["cora_ml", "citeseer"]


def test_chebyshev_step_count():
    import ppnp_b200 as P
    assert P.ppr_cheb_steps_for_tol(0.1, 1e-7) == 40 and P.ppr_steps_for_tol(0.1, 1e-7) == 153
    assert P.ppr_cheb_steps_for_tol(0.2, 1e-7) < P.ppr_cheb_steps_for_tol(0.1, 1e-7) < P.ppr_cheb_steps_for_tol(0.05, 1e-7)


def test_chebyshev_recurrence_reaches_the_reference_matrix_in_numpy():
    """The recurrence the kernel implements, in numpy fp32, against the reference's frozen Pi rows."""
    from util import oracle
    _, adj = load_std("citeseer")
    g = load_golden("citeseer")
    A = oracle.calc_A_hat(adj, "sym").astype(np.float32)
    n, alpha = A.shape[0], 0.1
    rho2 = (1 - alpha) ** 2
    I = np.eye(n, dtype=np.float32)
    prev, x, w = I, ((1 - alpha) * (A @ I) + alpha * I).astype(np.float32), 1.0
    for k in range(2, 41):
        w = 1 / (1 - rho2 / 2) if k == 2 else 1 / (1 - rho2 * w / 4)
        prev, x = x, (np.float32(w) * ((1 - alpha) * (A @ x) + alpha * I) + np.float32(1 - w) * prev).astype(np.float32)
    assert relerr(x[g["ppr_rows_idx"]], g["ppr_rows"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["sym", "rw"])
def test_ppr_dense_chebyshev_matches_reference_inverse(name, mode):
    import ppnp_b200 as P
    z, adj = load_std(name)
    dev = torch.device("cuda:0")
    ahat = P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev), torch.from_numpy(z["adj_indices"]).to(dev), None, mode)
    Pc = P.ppr_dense(ahat, 0.1, method="chebyshev")
    Pp = P.ppr_dense(ahat, 0.1)
    assert float((Pc - Pp).norm() / Pp.norm()) < 5e-6               # the same matrix as the plain iteration
    if mode == "sym":
        g = load_golden(name)
        got = Pc[torch.from_numpy(g["ppr_rows_idx"]).to(dev)].cpu().numpy()
        assert relerr(got, g["ppr_rows"]) < 1e-5                     # helpers.py:68-71, fp64 inverse (tolerances of
        assert relerr(torch.diagonal(Pc).cpu().numpy(), g["ppr_diag"]) < 1e-5   # test_ppr_dense_matches_reference_inverse)
        assert relerr(Pc.sum(1).cpu().numpy(), g["ppr_rowsum"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("K", [1, 2, 3, 6, 7])
def test_ppr_dense_chebyshev_small_K_matches_numpy_recurrence(K):
    """Odd and even step counts (the result must land in Pi either way), against the recurrence in fp64."""
    import ppnp_b200 as P
    from util import oracle
    z, adj = load_std("citeseer")
    dev = torch.device("cuda:0")
    ahat = P.csr_normalize(torch.from_numpy(z["adj_indptr"]).to(dev), torch.from_numpy(z["adj_indices"]).to(dev))
    A = oracle.calc_A_hat(adj, "sym").toarray()
    n, alpha = A.shape[0], 0.15
    rho2 = (1 - alpha) ** 2
    I = np.eye(n)
    prev, x, w = I, (1 - alpha) * A + alpha * I, 1.0
    for k in range(2, K + 1):
        w = 1 / (1 - rho2 / 2) if k == 2 else 1 / (1 - rho2 * w / 4)
        prev, x = x, w * ((1 - alpha) * (A @ x) + alpha * I) + (1 - w) * prev
    got = P.ppr_dense(ahat, alpha, K=K, method="chebyshev").cpu().numpy()
    assert relerr(got, x) < 1e-6
