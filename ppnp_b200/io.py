"""Graph files -> device CSR, without scipy in between.

``load_npz_graph`` reads the flat-dictionary ``.npz`` layout of the reference
(ppnp/data/sparsegraph.py:231-297: ``SparseGraph.to_flat_dict`` / ``from_flat_dict``; the files under
ppnp/data/ and what main.py:73-74 loads) straight into CUDA tensors; ``standardized_graph`` chains it with
the GPU standardisation (sparsegraph.py:191-222) and normalisation (helpers.py:58-66), i.e. everything
main.py:73-75 + helpers.calc_A_hat do on the host, and returns what the propagation consumes.

``save_csr_bin`` / ``load_csr_bin``: a flat binary CSR container for the synthetic graphs of the benchmark
configurations (an 8-word header, int64 row pointers, int32 columns), memory-mapped on load so that a
graph larger than host RAM's free part can still be streamed to the device in slabs.
"""
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

MAGIC = 0x50504E5043535231      # "PPNPCSR1"


@dataclass
class RawGraph:
    """Arrays of one SparseGraph file on the device: adjacency pattern (+ weights), attributes (CSR or
    dense), labels."""
    n: int
    adj_indptr: torch.Tensor            # int64 [n + 1]
    adj_indices: torch.Tensor           # int32 [nnz]
    adj_data: Optional[torch.Tensor]    # fp32 [nnz]
    attr_indptr: Optional[torch.Tensor] = None
    attr_indices: Optional[torch.Tensor] = None
    attr_data: Optional[torch.Tensor] = None
    attr_shape: Optional[tuple] = None
    attr_dense: Optional[torch.Tensor] = None
    labels: Optional[torch.Tensor] = None


def _matrix(z, name):
    """The four arrays of a sparse matrix in either separator convention of from_flat_dict
    (sparsegraph.py:255-276: '.' today, '_' and the short names 'adj' / 'attr' in older files)."""
    short = {"adj_matrix": "adj", "attr_matrix": "attr"}[name]
    for base, sep in ((name, "."), (name, "_"), (short, "."), (short, "_")):
        k = f"{base}{sep}data"
        if k in z:
            return (np.asarray(z[k]), np.asarray(z[f"{base}{sep}indices"]), np.asarray(z[f"{base}{sep}indptr"]),
                    tuple(int(x) for x in np.asarray(z[f"{base}{sep}shape"])))
    return None


def load_npz_graph(path, device="cuda"):
    """sparsegraph.py:247-297 ``SparseGraph.from_flat_dict(np.load(path))`` as device arrays."""
    dev = torch.device(device)
    with np.load(path, allow_pickle=True) as z:
        adj = _matrix(z, "adj_matrix")
        if adj is None:
            raise ValueError(f"{path}: no adjacency matrix (adj_matrix.data / .indices / .indptr / .shape)")
        data, indices, indptr, shape = adj
        if shape[0] != shape[1]:
            raise ValueError("Dimensions of the adjacency matrix don't agree.")      # sparsegraph.py:52-53
        if len(indptr) != shape[0] + 1 or int(indptr[-1]) != len(indices) or len(data) != len(indices):
            raise ValueError(f"{path}: inconsistent CSR arrays")
        g = RawGraph(n=shape[0],
                     adj_indptr=torch.from_numpy(indptr.astype(np.int64)).to(dev),
                     adj_indices=torch.from_numpy(indices.astype(np.int32)).to(dev),
                     adj_data=torch.from_numpy(data.astype(np.float32)).to(dev))
        attr = _matrix(z, "attr_matrix")
        if attr is not None:
            data, indices, indptr, shape = attr
            if shape[0] != g.n:
                raise ValueError("Dimensions of the adjacency and attribute matrices don't agree.")   # :66-67
            g.attr_indptr = torch.from_numpy(indptr.astype(np.int64)).to(dev)
            g.attr_indices = torch.from_numpy(indices.astype(np.int32)).to(dev)
            g.attr_data = torch.from_numpy(data.astype(np.float32)).to(dev)
            g.attr_shape = shape
        elif "attr_matrix" in z and z["attr_matrix"].dtype != object:
            g.attr_dense = torch.from_numpy(np.asarray(z["attr_matrix"], dtype=np.float32)).to(dev)
        if "labels" in z and z["labels"].dtype != object:
            lab = np.asarray(z["labels"])
            if lab.shape[0] != g.n:
                raise ValueError("Dimensions of the adjacency matrix and the label vector don't agree.")   # :76-77
            g.labels = torch.from_numpy(lab.astype(np.int64)).to(dev)
    return g


def standardized_graph(path_or_raw, device="cuda", mode="sym", select_lcc=True):
    """main.py:73-75 + helpers.py:58-66 on the device: file -> standardised adjacency -> A_hat.
    Returns (NormalizedCSR, keep, raw): ``keep`` = original ids of the kept nodes (subset attributes and
    labels with it, create_subgraph sparsegraph.py:345-351)."""
    from . import ops
    raw = load_npz_graph(path_or_raw, device) if isinstance(path_or_raw, (str, os.PathLike)) else path_or_raw
    ip, idx, keep = ops.graph_standardize(raw.adj_indptr, raw.adj_indices, select_lcc=select_lcc)
    return ops.csr_normalize(ip, idx, None, mode), keep, raw


def save_csr_bin(path, indptr, indices):
    """Flat binary CSR: 8 little-endian int64 words (magic, n, nnz, 0...), int64 indptr, int32 indices."""
    ip = np.ascontiguousarray(indptr.detach().cpu().numpy() if torch.is_tensor(indptr) else indptr, dtype="<i8")
    ix = np.ascontiguousarray(indices.detach().cpu().numpy() if torch.is_tensor(indices) else indices, dtype="<i4")
    n, nnz = len(ip) - 1, len(ix)
    if n < 0 or int(ip[-1]) != nnz:
        raise ValueError("indptr[-1] != len(indices)")
    head = np.zeros(8, dtype="<i8")
    head[0], head[1], head[2] = MAGIC, n, nnz
    with open(path, "wb") as f:
        f.write(head.tobytes())
        f.write(ip.tobytes())
        f.write(ix.tobytes())


def load_csr_bin(path, device="cuda", slab_bytes=256 << 20):
    """Memory-map the file and copy it to the device in slabs of ``slab_bytes`` (the host never holds a
    second copy).  Returns (indptr int64, indices int32) on ``device``."""
    dev = torch.device(device)
    head = np.fromfile(path, dtype="<i8", count=8)
    if len(head) < 8 or int(head[0]) != MAGIC:
        raise ValueError(f"{path}: not a ppnp_b200 CSR file")
    n, nnz = int(head[1]), int(head[2])
    want = 64 + 8 * (n + 1) + 4 * nnz
    if os.path.getsize(path) != want:
        raise ValueError(f"{path}: size {os.path.getsize(path)} != {want} expected from the header")
    ip_m = np.memmap(path, dtype="<i8", mode="r", offset=64, shape=(n + 1,))
    ix_m = np.memmap(path, dtype="<i4", mode="r", offset=64 + 8 * (n + 1), shape=(nnz,))

    def to_dev(m, dtype):
        out = torch.empty(len(m), dtype=dtype, device=dev)
        step = max(1, slab_bytes // m.dtype.itemsize)
        for a in range(0, len(m), step):
            out[a:a + step] = torch.from_numpy(np.ascontiguousarray(m[a:a + step])).to(dev)
        return out

    ip, ix = to_dev(ip_m, torch.int64), to_dev(ix_m, torch.int32)
    if int(ip[-1]) != nnz or int(ip[0]) != 0:
        raise ValueError(f"{path}: row pointers do not match the header")
    return ip, ix
