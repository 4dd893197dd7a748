"""Runs only the wide-output short-K GEMM (fwd L0 shape) a few times: target for `ncu --set full --import-source on`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taxoexpan_b200 import functional as txf
dev = torch.device("cuda", 0)
m, n, k = 37039, 2000, 300
A = torch.randn(m, txf.round8(k), device=dev)
B = torch.randn(n, txf.round8(k), device=dev) * 0.05
a16, b16 = txf.split_f16(A, k), txf.split_f16(B, k)
out = torch.empty(m, txf.round4(n), device=dev)
for _ in range(4):
    txf.gemm_nt_f16(a16, k, b16, n, out=out)
torch.cuda.synchronize()
