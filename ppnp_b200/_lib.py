"""ctypes binding of libppnp_b200.so (C ABI declared in include/ppnp_b200.h).

The library is the product: every op in this package calls it, and importing an op
without the built library raises -- there is no CPU or PyTorch fallback.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPNP_B200_LIB selects an experimental build of the same sources (tools/build_variants.py)
LIB_PATH = os.environ.get("PPNP_B200_LIB") or os.path.join(_HERE, "libppnp_b200.so")

# epilogues / modes (keep in sync with include/ppnp_b200.h)
MODE_SYM, MODE_RW, MODE_SYM_Y0 = 0, 1, 2
MODE_PER_STEP = 0x100      # PPNP_MODE_PER_STEP: never the one-launch cooperative kernel
EPI_PLAIN, EPI_Z2Y, EPI_Y, EPI_Y2Z, EPI_RW, EPI_Y02Z = 0, 1, 2, 3, 4, 5
EPI_ACC = 16
EPI_INPLACE = 32
STD_UNDIRECTED, STD_NO_SELF_LOOPS, STD_LCC = 1, 2, 4
FLAG = 0x80000000


class PlanStruct(C.Structure):
    """Mirror of ppnp_plan_t."""
    _fields_ = [
        ("n", C.c_int64), ("n_edges", C.c_int64), ("n_chunks", C.c_int64), ("n_segs", C.c_int64),
        ("n_fix", C.c_int64), ("n_slots", C.c_int64), ("chunk_edges", C.c_int32), ("flags", C.c_int32),
        ("cols", C.c_void_p), ("vals", C.c_void_p), ("seg_row", C.c_void_p), ("chunk_seg", C.c_void_p),
        ("fix_ptr", C.c_void_p), ("fix_row", C.c_void_p), ("fix_deg", C.c_void_p), ("row_deg", C.c_void_p),
    ]


class TiledPlanStruct(C.Structure):
    """Mirror of ppnp_tiled_plan_t."""
    _fields_ = [
        ("n", C.c_int64), ("n_slabs", C.c_int64), ("n_pieces", C.c_int64), ("n_ctas", C.c_int32),
        ("warps_per_cta", C.c_int32), ("slots_cap", C.c_int32), ("slack", C.c_int32),
        ("cols", C.c_void_p), ("vals", C.c_void_p), ("slab_meta", C.c_void_p), ("piece_slot", C.c_void_p),
        ("warp_slab_ptr", C.c_void_p), ("cta_slot_ptr", C.c_void_p), ("slot_row", C.c_void_p), ("row_deg", C.c_void_p),
    ]


class RowsPlanStruct(C.Structure):
    """Mirror of ppnp_rows_plan_t."""
    _fields_ = [("n", C.c_int64), ("n_rows", C.c_int64), ("indptr", C.c_void_p), ("indices", C.c_void_p),
                ("vals", C.c_void_p), ("rows", C.c_void_p)]


_p, _i32, _i64, _f32, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64

# name -> (restype, argtypes).  Every symbol include/ppnp_b200.h declares is listed here;
# tests/test_abi.py checks the two stay in sync.
SIGNATURES = {
    "ppnp_last_error": (C.c_char_p, []),
    "ppnp_version": (C.c_int, []),
    "ppnp_device_info": (C.c_int, [_p, _p, _p, _p]),
    "ppnp_csr_normalize_workspace_bytes": (_i64, [_i64]),
    "ppnp_csr_normalize": (C.c_int, [_p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ppnp_spmm_step": (C.c_int, [C.POINTER(PlanStruct), _p, _p, _p, _p, _i64, _i32, _f32, _i32, _i32, _p]),
    "ppnp_spmm_step_push": (C.c_int, [C.POINTER(PlanStruct), _p, _p, _p, _p, _i64, _i32, _f32, _i32, _i32, _p, _p, _p, _p, _i32, _p]),
    "ppnp_appnp_propagate": (C.c_int, [C.POINTER(PlanStruct), _p, _p, _p, _p, _i64, _i32, _i32, _f32, _i32, _i32, _p]),
    "ppnp_appnp_propagate_persistent": (C.c_int, [C.POINTER(PlanStruct), _p, _p, _p, _p, _i64, _i32, _i32, _f32, _p]),
    "ppnp_spmm_step_tiled": (C.c_int, [C.POINTER(TiledPlanStruct), _p, _p, _p, _i64, _i32, _i32, _f32, _i32, _i32, _p]),
    "ppnp_spmm_step_rows": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _p, _p, _p, _i64, _i32, _f32, _i32, _i32, _p, _p, _p, _p, _i32, _p]),
    "ppnp_appnp_propagate_parts": (C.c_int, [C.POINTER(TiledPlanStruct), C.POINTER(PlanStruct), C.POINTER(RowsPlanStruct), _p, _p, _p, _p,
                                             _i64, _i32, _i32, _i32, _f32, _i32, _i32, _p]),
    "ppnp_linear_rowscale": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _i64, _i32, _p]),
    "ppnp_linear_rowscale_backward_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "ppnp_linear_rowscale_backward": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ppnp_ppr_dense": (C.c_int, [_p, _p, _p, _i64, _f32, _i32, _p, _p, _p]),
    "ppnp_ppr_dense_cheb": (C.c_int, [_p, _p, _p, _i64, _f32, _i32, _p, _p, _p]),
    "ppnp_gather_gemm_f32": (C.c_int, [_p, _i64, _p, _i64, _i64, _p, _i64, _i32, _p, _i64, _i32, _p]),
    "ppnp_gather_gemm_bf16_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "ppnp_gather_gemm_bf16": (C.c_int, [_p, _i64, _p, _i64, _i64, _p, _i64, _i32, _p, _i64, _p, _i64, _p]),
    "ppnp_f32_to_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "ppnp_topk_thresh": (C.c_int, [_p, _i64, _i64, _i64, _i32, _p, _p]),
    "ppnp_topk_mask": (C.c_int, [_p, _i64, _i64, _i64, _p, _p]),
    "ppnp_dense_row_nnz": (C.c_int, [_p, _i64, _i64, _i64, _p, _p]),
    "ppnp_dense_to_csr": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "ppnp_batch_support": (C.c_int, [_p, _p, _p, _i64, _p, _p]),
    "ppnp_batch_support_colmap": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "ppnp_batch_propagate": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p, _i64, _i32, _p, _i64, _i32, _p]),
    "ppnp_plan_workspace_bytes": (_i64, [_i64]),
    "ppnp_plan_measure": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ppnp_plan_fill": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64,
                                 _p, _p, _p, _p, _p, _p, _p, _p]),
    "ppnp_gather_rows": (C.c_int, [_p, _i64, _p, _i64, _i32, _p, _i64, _p]),
    "ppnp_rmat_keys": (C.c_int, [_u64, _i32, _i64, _i64, _i64, _p, _p]),
    "ppnp_graph_standardize_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "ppnp_graph_standardize": (C.c_int, [_p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _i64, _p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"ppnp_b200: {LIB_PATH} is missing -- the CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` in the repo root. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = ABI drift, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ppnp_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
