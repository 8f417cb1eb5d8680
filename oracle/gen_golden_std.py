#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- freeze inputs and outputs of the REFERENCE's graph standardisation
(ppnp/data/sparsegraph.py:191-222 SparseGraph.standardize and what it calls: to_unweighted
:150-154, to_undirected :127-148, remove_self_loops :381-395, largest_connected_components
:355-379, create_subgraph :300-352) by importing and running it unmodified:

    python oracle/gen_golden_std.py      (build container only; /root/reference is read-only)

Writes tests/golden/standardize_cases.npz: for every case the raw CSR the reference was given and
the CSR + kept node ids it returned.  Cases: the two data sets the reference ships (raw files,
main.py:73-75) and small synthetic graphs that exercise what the shipped files do not (weights,
one-directional edges, self loops, several components, isolated nodes, flag combinations).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

REF = "/root/reference"
sys.path.insert(0, REF)
from ppnp.data.sparsegraph import SparseGraph  # noqa: E402  (reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "standardize_cases.npz")


def run_reference(adj, **flags):
    n = adj.shape[0]
    g = SparseGraph(adj_matrix=adj.copy(), node_names=np.arange(n))
    g = g.standardize(**flags)
    out = g.adj_matrix.tocsr()
    out.sort_indices()
    assert out.has_canonical_format
    return out, np.asarray(g.node_names, dtype=np.int64)


def random_graph(rng, n, m, weighted, loops, blocks):
    """m directed edges inside `blocks` disjoint node groups (=> several components)."""
    size = n // blocks
    b = rng.randint(0, blocks, m)
    r = b * size + rng.randint(0, size, m)
    c = b * size + rng.randint(0, size, m)
    if not loops:
        keep = r != c
        r, c = r[keep], c[keep]
    w = rng.randint(1, 4, len(r)).astype(np.float32) if weighted else np.ones(len(r), dtype=np.float32)
    a = sp.csr_matrix((w, (r, c)), shape=(n, n))
    a.sum_duplicates()
    if weighted:       # the reference insists that opposing edges carry equal weights (sparsegraph.py:138-139)
        a = a.maximum(a.T).tocsr() if rng.rand() < 0.5 else a
    a.data = np.where(a.data > 0, a.data, 1).astype(np.float32)
    a.sort_indices()
    return a


def main():
    cases = {}
    flags_all = dict(make_unweighted=True, make_undirected=True, no_self_loops=True, select_lcc=True)
    for name in ("cora_ml", "citeseer"):
        raw = np.load(os.path.join(REF, "ppnp", "data", f"{name}.npz"), allow_pickle=True)
        adj = sp.csr_matrix((raw["adj_matrix.data"], raw["adj_matrix.indices"], raw["adj_matrix.indptr"]),
                            shape=tuple(raw["adj_matrix.shape"]))
        cases[name] = (adj, flags_all)
    rng = np.random.RandomState(0)
    cases["directed_loops_3blocks"] = (random_graph(rng, 300, 900, False, True, 3), flags_all)
    cases["weighted_loops_5blocks"] = (random_graph(rng, 500, 1500, True, True, 5), flags_all)
    cases["sparse_many_isolated"] = (random_graph(rng, 1000, 700, False, True, 1), flags_all)
    cases["keep_loops"] = (random_graph(rng, 200, 800, False, True, 2), dict(flags_all, no_self_loops=False))
    cases["no_lcc"] = (random_graph(rng, 200, 300, False, True, 4), dict(flags_all, select_lcc=False))
    cases["directed_lcc"] = (random_graph(rng, 240, 700, False, False, 3), dict(flags_all, make_undirected=False))
    # two components of equal size: which one the reference keeps is whatever np.argsort(...)[::-1][0] says
    r = np.array([0, 1, 2, 3, 4, 5]); c = np.array([1, 2, 0, 4, 5, 3])
    cases["tie_two_triangles"] = (sp.csr_matrix((np.ones(6, dtype=np.float32), (r, c)), shape=(6, 6)), flags_all)
    cases["already_standard"] = (run_reference(cases["directed_loops_3blocks"][0], **flags_all)[0], flags_all)

    out = {"names": np.array(sorted(cases))}
    for name, (adj, flags) in cases.items():
        adj = adj.tocsr().astype(np.float32)
        adj.sort_indices()
        res, keep = run_reference(adj, **flags)
        out[f"{name}.in_indptr"] = adj.indptr.astype(np.int64)
        out[f"{name}.in_indices"] = adj.indices.astype(np.int32)
        out[f"{name}.in_data"] = adj.data.astype(np.float32)
        out[f"{name}.flags"] = np.array([flags["make_unweighted"], flags["make_undirected"], flags["no_self_loops"],
                                         flags["select_lcc"]], dtype=np.int32)
        out[f"{name}.out_indptr"] = res.indptr.astype(np.int64)
        out[f"{name}.out_indices"] = res.indices.astype(np.int32)
        out[f"{name}.out_data"] = res.data.astype(np.float32)
        out[f"{name}.keep"] = keep
        print(f"{name:26s} n {adj.shape[0]:5d} nnz {adj.nnz:6d} -> n {res.shape[0]:5d} nnz {res.nnz:6d} "
              f"data all ones: {bool((res.data == 1).all())}")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
