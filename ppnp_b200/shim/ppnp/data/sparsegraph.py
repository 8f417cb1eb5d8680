"""Drop-in for the reference's ``ppnp.data.sparsegraph`` (main.py:27, batch-main.py:28): the same module
with ``SparseGraph.standardize`` (sparsegraph.py:191-222) running on the GPU.

The reference's own file (the next ``ppnp/data/sparsegraph.py`` on sys.path) is loaded under a
private name and re-exported unchanged -- container class, npz layout (from_flat_dict / to_flat_dict),
create_subgraph, ...; only ``standardize`` is replaced: symmetrise, drop self loops, largest connected
component and compaction are one C-ABI call (``ppnp_graph_standardize``, csrc/standardize.cu) whose
CSR is bit-exact with the reference's.  There is no CPU fallback: without CUDA it raises.
"""
import importlib.util
import os
import sys

import numpy as np
import scipy.sparse as sp


def _load_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    for d in sys.path:
        cand = os.path.join(os.path.abspath(d or "."), "ppnp", "data", "sparsegraph.py")
        if os.path.isfile(cand) and os.path.dirname(cand) != here:
            spec = importlib.util.spec_from_file_location("_ppnp_reference_sparsegraph", cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("the reference's ppnp/data/sparsegraph.py must be on sys.path behind the shim directory")


_ref = _load_reference()
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})


class SparseGraph(_ref.SparseGraph):
    def standardize(self, make_unweighted=True, make_undirected=True, no_self_loops=True, select_lcc=True):
        """sparsegraph.py:191-222, in place, same return value (self)."""
        import torch

        import ppnp_b200 as P
        if not make_unweighted:
            raise NotImplementedError("the GPU standardisation is the unit-weight pipeline of main.py:75")
        if not torch.cuda.is_available():
            raise RuntimeError("ppnp_b200: SparseGraph.standardize needs a CUDA device (there is no CPU path)")
        adj = self.adj_matrix.tocsr()
        if adj.nnz and (adj.data == 0).any():
            adj.eliminate_zeros()
        n = adj.shape[0]
        dev = torch.device("cuda")
        ip, idx, keep = P.graph_standardize(torch.from_numpy(adj.indptr.astype(np.int64)).to(dev),
                                            torch.from_numpy(adj.indices.astype(np.int32)).to(dev),
                                            True, make_undirected, no_self_loops, select_lcc)
        ip, idx, keep = ip.cpu().numpy(), idx.cpu().numpy(), keep.cpu().numpy()
        m = len(keep)
        self.adj_matrix = sp.csr_matrix((np.ones(len(idx), dtype=np.float32), idx, ip), shape=(m, m))
        if m != n:      # create_subgraph, sparsegraph.py:345-351
            if self.attr_matrix is not None:
                self.attr_matrix = self.attr_matrix[keep]
            if self.labels is not None:
                self.labels = self.labels[keep]
            if self.node_names is not None:
                self.node_names = self.node_names[keep]
        return self


# from_flat_dict (sparsegraph.py:247-297) builds ``SparseGraph(**init_dict)`` through its module's globals
_ref.SparseGraph = SparseGraph
