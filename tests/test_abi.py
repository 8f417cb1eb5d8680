"""The C-ABI library loads and exports every symbol include/ppnp_b200.h declares (CPU only: no
compute calls are made here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ppnp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ppnp_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["ppnp_csr_normalize", "ppnp_appnp_propagate", "ppnp_spmm_step", "ppnp_ppr_dense",
              "ppnp_gather_gemm_f32", "ppnp_gather_gemm_bf16", "ppnp_topk_thresh", "ppnp_topk_mask",
              "ppnp_batch_support", "ppnp_batch_propagate", "ppnp_last_error"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from ppnp_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/ppnp_b200.h but not exported"


def test_binding_table_matches_header():
    from ppnp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.ppnp_version() >= 100
    assert isinstance(lib.ppnp_last_error(), bytes)


def test_plan_struct_layout_matches_header():
    from ppnp_b200 import _lib
    # 6 x int64, 2 x int32, 8 pointers
    assert ctypes.sizeof(_lib.PlanStruct) == 6 * 8 + 2 * 4 + 8 * 8


def test_missing_library_fails_loudly(monkeypatch):
    from ppnp_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libppnp_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_ops_refuse_cpu_tensors():
    import torch
    import ppnp_b200
    with pytest.raises(RuntimeError, match="no CPU path"):
        ppnp_b200.csr_normalize(torch.zeros(3, dtype=torch.int32), torch.zeros(2, dtype=torch.int32))
