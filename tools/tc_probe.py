#!/usr/bin/env python
"""Small driver for an ncu capture of the tcgen05 gather-GEMM at BASELINE config 2 shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppnp_b200 as P
n = 19717
dev = torch.device("cuda:0")
Pi = torch.rand(n, n, device=dev)
Pb = P.to_bf16_padded(Pi)
H = torch.randn(n, 7, device=dev)
for _ in range(3):
    out = P.gather_gemm_bf16(Pb, H, None)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
