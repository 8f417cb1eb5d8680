"""Graph files -> arrays (ppnp_b200/io.py): the reference's flat-dict npz layout
(sparsegraph.py:231-297) and the flat binary CSR container.  CPU: parsing and round trips; GPU
(marked): file -> standardise -> normalise equals the reference's calc_A_hat on its own pipeline."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, load_golden, load_std
from ppnp_b200 import io as pio

CASES = np.load(os.path.join(GOLDEN, "standardize_cases.npz"))


def write_flat_dict(path, name, sep=".", short=False, with_attr=True):
    """A file in to_flat_dict's layout (sparsegraph.py:231-245) from the frozen raw arrays."""
    ip, idx, data = CASES[f"{name}.in_indptr"], CASES[f"{name}.in_indices"], CASES[f"{name}.in_data"]
    n = len(ip) - 1
    a = "adj" if short else "adj_matrix"
    d = {f"{a}{sep}data": data, f"{a}{sep}indices": idx, f"{a}{sep}indptr": ip.astype(np.int32), f"{a}{sep}shape": np.array([n, n]),
         "labels": np.arange(n) % 7, "node_names": np.array(None, dtype=object), "metadata": np.array(None, dtype=object)}
    if with_attr:
        t = "attr" if short else "attr_matrix"
        d.update({f"{t}{sep}data": np.ones(n, dtype=np.float32), f"{t}{sep}indices": (np.arange(n) % 5).astype(np.int32),
                  f"{t}{sep}indptr": np.arange(n + 1, dtype=np.int32), f"{t}{sep}shape": np.array([n, 5])})
    np.savez(path, **d)
    return n


@pytest.mark.parametrize("sep,short", [(".", False), ("_", False), ("_", True)])
def test_load_npz_graph_layouts(tmp_path, sep, short):
    p = str(tmp_path / "g.npz")
    n = write_flat_dict(p, "citeseer", sep, short)
    g = pio.load_npz_graph(p, device="cpu")
    assert g.n == n == 3312 and g.adj_indptr.dtype == torch.int64 and g.adj_indices.dtype == torch.int32
    assert np.array_equal(g.adj_indptr.numpy(), CASES["citeseer.in_indptr"])
    assert np.array_equal(g.adj_indices.numpy(), CASES["citeseer.in_indices"])
    assert np.array_equal(g.adj_data.numpy(), CASES["citeseer.in_data"])
    assert g.attr_shape == (n, 5) and g.attr_indices.numel() == n and g.labels.tolist()[:8] == [0, 1, 2, 3, 4, 5, 6, 0]


def test_load_npz_graph_errors(tmp_path):
    p = str(tmp_path / "bad.npz")
    np.savez(p, labels=np.arange(3))
    with pytest.raises(ValueError, match="no adjacency"):
        pio.load_npz_graph(p, device="cpu")
    np.savez(p, **{"adj_matrix.data": np.ones(1, np.float32), "adj_matrix.indices": np.zeros(1, np.int32),
                   "adj_matrix.indptr": np.array([0, 1, 1]), "adj_matrix.shape": np.array([2, 3])})
    with pytest.raises(ValueError, match="don't agree"):
        pio.load_npz_graph(p, device="cpu")
    np.savez(p, **{"adj_matrix.data": np.ones(1, np.float32), "adj_matrix.indices": np.zeros(1, np.int32),
                   "adj_matrix.indptr": np.array([0, 1, 2]), "adj_matrix.shape": np.array([2, 2])})
    with pytest.raises(ValueError, match="inconsistent"):
        pio.load_npz_graph(p, device="cpu")


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("name", ["cora_ml", "citeseer"])
def test_load_npz_graph_reads_the_reference_files(name):
    g = pio.load_npz_graph(f"/root/reference/ppnp/data/{name}.npz", device="cpu")
    assert np.array_equal(g.adj_indptr.numpy(), CASES[f"{name}.in_indptr"])
    # the raw citeseer file stores its rows unsorted; the frozen case holds them sorted
    rows = np.repeat(np.arange(g.n), np.diff(g.adj_indptr.numpy()))
    key = np.sort(rows * g.n + g.adj_indices.numpy())
    assert np.array_equal(key, rows * g.n + CASES[f"{name}.in_indices"])
    assert g.labels is not None and g.attr_shape[0] == g.n


def test_csr_bin_round_trip_and_validation(tmp_path):
    z, _ = load_std("cora_ml")
    p = str(tmp_path / "g.csr")
    pio.save_csr_bin(p, z["adj_indptr"], z["adj_indices"])
    assert os.path.getsize(p) == 64 + 8 * len(z["adj_indptr"]) + 4 * len(z["adj_indices"])
    ip, ix = pio.load_csr_bin(p, device="cpu", slab_bytes=4096)         # many slabs
    assert np.array_equal(ip.numpy(), z["adj_indptr"]) and np.array_equal(ix.numpy(), z["adj_indices"])
    with open(p, "r+b") as f:
        f.truncate(os.path.getsize(p) - 4)
    with pytest.raises(ValueError, match="size"):
        pio.load_csr_bin(p, device="cpu")
    with open(p, "wb") as f:
        f.write(b"\0" * 128)
    with pytest.raises(ValueError, match="not a ppnp_b200"):
        pio.load_csr_bin(p, device="cpu")
    with pytest.raises(ValueError):
        pio.save_csr_bin(p, np.array([0, 2]), np.array([1]))
