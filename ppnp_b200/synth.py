"""Synthetic graphs for BASELINE.json configs 2-5 (benchmark / test inputs, not reference code).

R-MAT recipe of SURVEY.md section 8(d): (a, b, c, d) = (0.57, 0.19, 0.19, 0.05), ids truncated to n,
loops dropped, symmetrised, de-duplicated -> the canonical CSR that ``SparseGraph.standardize``
would hand to the hot path (sorted int32 indices, unit weights, zero diagonal).  Edges come from
the counter-based generator in include/ppnp_rmat.h (device kernel ppnp_rmat_keys); sorting and
de-duplication use torch.sort / unique as plumbing.
"""
import torch

from . import _lib


def rmat_adjacency(n, raw_draws, scale, seed=0, device="cuda", e0=0, row_range=None):
    """Returns (indptr int64 [rows+1], indices int32 [nnz]) of the symmetrised R-MAT graph.
    ``row_range=(lo, hi)`` keeps only rows lo..hi-1 (a shard of the 1-D row partition)."""
    lib = _lib.load()
    dev = torch.device(device)
    keys = torch.empty(2 * raw_draws, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.ppnp_rmat_keys(int(seed), int(scale), int(n), int(e0), int(e0 + raw_draws), _lib.ptr(keys),
                                _lib.current_stream())
    _lib.check(rc, "ppnp_rmat_keys")
    keys = keys[keys >= 0]
    if row_range is not None:
        lo, hi = row_range
        keys = keys[(keys >= (lo << 32)) & (keys < (hi << 32))]
    keys = torch.unique(keys, sorted=True)
    rows = keys >> 32
    indices = (keys & 0xFFFFFFFF).to(torch.int32)
    lo, hi = (0, n) if row_range is None else row_range
    counts = torch.bincount(rows - lo, minlength=hi - lo)
    indptr = torch.zeros(hi - lo + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, indices


def powerlaw_adjacency(n, n_edges_target, seed=0, device="cuda", exponent=2.2):
    """Degree-skewed random graph of PubMed shape (config 2/3: n = 19 717, nnz(A) ~ 88 648):
    endpoints drawn from a Zipf-like weight w_i ~ (i + 10)^(-1/(exponent-1)), symmetrised."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = (torch.arange(n, dtype=torch.float64) + 10.0) ** (-1.0 / (exponent - 1.0))
    m = int(n_edges_target * 0.56)
    src = torch.multinomial(w, m, replacement=True, generator=g)
    dst = torch.multinomial(w, m, replacement=True, generator=g)
    perm = torch.randperm(n, generator=g)
    src, dst = perm[src], perm[dst]
    # a ring keeps the graph connected like the reference's LCC inputs
    ring = torch.arange(n)
    src = torch.cat([src, ring])
    dst = torch.cat([dst, (ring + 1) % n])
    keep = src != dst
    src, dst = src[keep], dst[keep]
    keys = torch.cat([(src << 32) | dst, (dst << 32) | src])
    keys = torch.unique(keys, sorted=True)
    rows = keys >> 32
    indices = (keys & 0xFFFFFFFF).to(torch.int32)
    counts = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr.to(device), indices.to(device)
