"""Graph standardisation: the numpy restatement in oracle/ against the
reference's own outputs frozen by oracle/gen_golden_std.py, and its agreement with the hot path's
input fixtures (tests/golden/*_std.npz, produced by the reference through main.py:73-75)."""
import os

import numpy as np
import pytest

from util import GOLDEN, load_std, oracle

CASES = np.load(os.path.join(GOLDEN, "standardize_cases.npz"))
NAMES = [str(x) for x in CASES["names"]]


def case(name):
    f = CASES[f"{name}.flags"]
    flags = dict(make_unweighted=bool(f[0]), make_undirected=bool(f[1]), no_self_loops=bool(f[2]), select_lcc=bool(f[3]))
    return (CASES[f"{name}.in_indptr"], CASES[f"{name}.in_indices"], CASES[f"{name}.in_data"], flags,
            CASES[f"{name}.out_indptr"], CASES[f"{name}.out_indices"], CASES[f"{name}.keep"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_standardize_matches_reference(name):
    ip, idx, data, flags, oip, oidx, keep = case(name)
    got_ip, got_idx, got_keep = oracle.standardize(ip, idx, data, **flags)
    assert np.array_equal(got_ip, oip) and np.array_equal(got_idx, oidx) and np.array_equal(got_keep, keep)
    assert (CASES[f"{name}.out_data"] == 1).all()


@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
def test_standardized_fixture_is_the_hot_path_input(name):
    """The frozen output for the raw data sets IS the adjacency the propagation tests consume."""
    z, _ = load_std(name)
    assert np.array_equal(CASES[f"{name}.out_indptr"], z["adj_indptr"])
    assert np.array_equal(CASES[f"{name}.out_indices"], z["adj_indices"])


def test_tie_between_largest_components_follows_argsort():
    ip, idx, data, flags, oip, oidx, keep = case("tie_two_triangles")
    assert keep.tolist() == [3, 4, 5]       # np.argsort([3, 3])[::-1][0] == 1: the component with the larger smallest node


def test_standardize_is_idempotent_and_rejects_unsupported():
    ip, idx, data, flags, oip, oidx, keep = case("directed_loops_3blocks")
    a = oracle.standardize(ip, idx, data)
    b = oracle.standardize(a[0], a[1], None)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(b[2], np.arange(len(a[2])))
    with pytest.raises(NotImplementedError):
        oracle.standardize(ip, idx, data, make_unweighted=False)
    with pytest.raises(ValueError):
        oracle.standardize(ip, idx, np.zeros_like(data))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_oracle_standardize_matches_reference_on_random_graphs():
    """Beyond the frozen cases: 40 random graphs (directed, weighted, loops, several components, isolated
    nodes, every flag combination the unit-weight pipeline has) through the reference itself."""
    import sys
    import warnings
    import scipy.sparse as sp
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_sparsegraph_for_tests", "/root/reference/ppnp/data/sparsegraph.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # by path: independent of whatever `ppnp` package other tests imported
    SparseGraph = mod.SparseGraph
    rng = np.random.RandomState(11)
    for trial in range(40):
        n = int(rng.randint(2, 120))
        m = int(rng.randint(0, 4 * n))
        blocks = int(rng.randint(1, 5))
        size = max(1, n // blocks)
        b = rng.randint(0, blocks, m)
        r = np.minimum(b * size + rng.randint(0, size, m), n - 1)
        c = np.minimum(b * size + rng.randint(0, size, m), n - 1)
        a = sp.csr_matrix((np.ones(m, dtype=np.float32), (r, c)), shape=(n, n))
        a.sum_duplicates()
        a.data[:] = 1.0
        a.sort_indices()
        flags = dict(make_unweighted=True, make_undirected=bool(rng.rand() < 0.8), no_self_loops=bool(rng.rand() < 0.8),
                     select_lcc=bool(rng.rand() < 0.8))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g = SparseGraph(adj_matrix=a.copy(), node_names=np.arange(n)).standardize(**flags)
        ref = g.adj_matrix.tocsr()
        ref.sort_indices()
        sizes_tie = False
        if flags["select_lcc"]:
            # the reference's choice between equally large components is np.argsort's (unstable above 16 elements):
            # compare only when the largest component is unique
            comp = sp.csgraph.connected_components(a + a.T if True else a)[1]
            cnt = np.bincount(comp)
            sizes_tie = (cnt == cnt.max()).sum() > 1
        if sizes_tie:
            continue
        ip, idx, keep = oracle.standardize(a.indptr, a.indices, a.data, **flags)
        assert np.array_equal(ip, ref.indptr) and np.array_equal(idx, ref.indices), (trial, flags)
        assert np.array_equal(keep, np.asarray(g.node_names)), (trial, flags)
