"""ppnp_b200 -- the PPNP/APPNP propagation hot path of bkj/ppnp as hand-written sm_100a CUDA
behind the reference's own ``model.PPNP`` / ``helpers.compute_ppr`` interface.

Layout: csrc/ (CUDA kernels + C ABI, built into libppnp_b200.so), _lib.py (ctypes binding),
plan.py (edge-stream bookkeeping), ops.py (tensor-level operators), dist.py (row-partitioned
multi-GPU propagation), shim/ (drop-in ``model`` and ``helpers`` modules), synth.py (synthetic
R-MAT graphs for the benchmark configurations).
"""
from . import _lib  # noqa: F401
from .ops import (NormalizedCSR, PropagationGraph, SparseInput, appnp_fused_tail, linear_rowscale, linear_rowscale_backward, graph_standardize, sparse_first_layer, SparsePPR, appnp, appnp_propagate, appnp_propagate_persistent, batch_propagate,  # noqa: F401
                  batch_support, csr_normalize, dense_to_sparse_ppr, gather_gemm, gather_rows, gather_gemm_bf16, ppr_dense,
                  ppr_matmul, ppr_steps_for_tol, ppr_cheb_steps_for_tol, spmm_step, to_bf16, to_bf16_padded, topk_sparsify_, topk_thresh)
from .plan import StreamPlan, build_carved_plan, build_stream_plan, degree_order, lane_transpose  # noqa: F401
from .tiled import TiledPlan, build_tiled_plan  # noqa: F401

__version__ = "0.1.0"
