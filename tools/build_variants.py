#!/usr/bin/env python
"""Build experimental variants of libppnp_b200.so (same sources, different tuning macros) next to
the default library.  Select one at run time with PPNP_B200_LIB=<path>."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppnp_b200 import build as b  # noqa: E402

VARIANTS = {
    "mb3_u4": ["PPNP_SPMM_MINBLOCKS=3", "PPNP_SPMM_U4=4"],
    "mb3_u8": ["PPNP_SPMM_MINBLOCKS=3", "PPNP_SPMM_U4=8"],
    "mb2_u8": ["PPNP_SPMM_MINBLOCKS=2", "PPNP_SPMM_U4=8"],
    "mb4_u2": ["PPNP_SPMM_MINBLOCKS=4", "PPNP_SPMM_U4=2"],
    # slabs with <= 2 segment ends per lane group keep the rolling gather ring (profiles/r01_variants.md, point 3)
    "fewends": ["PPNP_SPMM_FEWENDS=1"],
    "fewends_segpred": ["PPNP_SPMM_FEWENDS=1", "PPNP_SPMM_SEGPRED=1"],
    # rows per CTA of the dense-Pi build step (csrc/ppr_dense.cu)
    "pprrows1": ["PPNP_PPR_ROWS=1"],
    "pprrows4": ["PPNP_PPR_ROWS=4"],
    "pprrows16": ["PPNP_PPR_ROWS=16"],
    "pprrows32": ["PPNP_PPR_ROWS=32"],
}

if __name__ == "__main__":
    outdir = os.path.join(ROOT, "ppnp_b200", "variants")
    os.makedirs(outdir, exist_ok=True)
    for name in (sys.argv[1:] or VARIANTS):
        out = os.path.join(outdir, f"libppnp_b200_{name}.so")
        b.build_library(defines=VARIANTS[name], out=out)
        print(out)
