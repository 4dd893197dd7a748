"""Time the tf32x3 and f16x3 GEMMs at the MAG-CS layer shapes (CUDA events, inputs rotated to defeat L2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taxoexpan_b200 import functional as txf
dev = torch.device("cuda", 0)
N = 37039
def timeit(fn, reps=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
NB = 3
for name, (m, n, k) in {"fwd L0": (N, 2000, 300), "fwd L1": (N, 500, 2050), "dz L1": (N, 2050, 500), "dz L0": (N, 52, 2000)}.items():
    As = [torch.randn(m, txf.round8(k), device=dev) for _ in range(NB)]
    B = torch.randn(n, txf.round8(k), device=dev) * 0.05
    t32 = [txf.split_tf32(a, k) for a in As]; b32 = txf.split_tf32(B, k)
    t16 = [txf.split_f16(a, k) for a in As]; b16 = txf.split_f16(B, k)
    out = torch.empty(m, txf.round4(n), device=dev)
    ms32 = timeit(lambda i: txf.gemm_nt_ps(t32[i % NB][0], t32[i % NB][1], k, b32[0], b32[1], n, out=out))
    ms16 = timeit(lambda i: txf.gemm_nt_f16(t16[i % NB], k, b16, n, out=out))
    fl = 2.0 * m * n * k
    print(f"NT {name:7s} {m}x{n}x{k}: tf32x3 {ms32:.4f} ms ({fl/ms32/1e9:.0f} TF/s eff)   f16x3 {ms16:.4f} ms ({fl/ms16/1e9:.0f} TF/s eff)   x{ms32/ms16:.2f}")
for name, (r, m, n) in {"dW L0": (N, 2000, 300), "dW L1": (N, 500, 2050)}.items():
    As = [torch.randn(r, txf.round8(m), device=dev) * 1e-3 for _ in range(NB)]
    Bs = [torch.randn(r, txf.round8(n), device=dev) for _ in range(NB)]
    a32 = [txf.split_tf32(a, m) for a in As]; b32 = [txf.split_tf32(b, n) for b in Bs]
    a16 = [txf.split_f16(a, m) for a in As]; b16 = [txf.split_f16(b, n) for b in Bs]
    ms32 = timeit(lambda i: txf.gemm_tn_ps(a32[i % NB][0], a32[i % NB][1], m, b32[i % NB][0], b32[i % NB][1], n))
    ms16 = timeit(lambda i: txf.gemm_tn_f16(a16[i % NB], m, b16[i % NB], n))
    fl = 2.0 * r * m * n
    print(f"TN {name:7s} r={r} {m}x{n}: tf32x3 {ms32:.4f} ms ({fl/ms32/1e9:.0f} TF/s eff)   f16x3 {ms16:.4f} ms ({fl/ms16/1e9:.0f} TF/s eff)   x{ms32/ms16:.2f}")
