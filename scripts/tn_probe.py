"""Debug probe (not a test): error of the MN-major GEMM for one descriptor variant given by TAXO_TN_* env vars."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from taxoexpan_b200 import functional as txf
dev = torch.device("cuda", 0)
for (r, m, n) in [(64, 128, 64), (256, 128, 64), (1000, 500, 300)]:
    g = torch.Generator().manual_seed(1)
    a = torch.randn(r, m, generator=g) / np.sqrt(r)
    b = torch.randn(r, n, generator=g)
    ref = a.double().t() @ b.double()
    a_hi, a_lo = txf.split_tf32(a.to(dev)); b_hi, b_lo = txf.split_tf32(b.to(dev))
    got = txf.gemm_tn_ps(a_hi, a_lo, m, b_hi, b_lo, n).cpu().double()
    print({k: os.environ.get(k) for k in ("TAXO_TN_SWIZZLE", "TAXO_TN_LAYOUT", "TAXO_TN_LBO", "TAXO_TN_SBO")}, (r, m, n),
          "err %.3e" % float((got - ref).abs().max()), "ref %.3f" % float(ref.abs().max()))
