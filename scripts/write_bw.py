"""HBM write ceiling: time torch fill_ / copy_ on the GEMM's 296 MB output (CUDA events)."""
import torch
dev = torch.device("cuda", 0)
outs = [torch.empty(37039, 2000, device=dev) for _ in range(4)]
src = torch.randn(37039, 2000, device=dev)
def t(fn, reps=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
b = outs[0].numel() * 4
ms = t(lambda i: outs[i % 4].fill_(1.0)); print(f"fill_ 296 MB: {ms:.4f} ms  {b/ms/1e6:.0f} GB/s write")
ms = t(lambda i: outs[i % 4].copy_(src)); print(f"copy_ 296 MB: {ms:.4f} ms  {2*b/ms/1e6:.0f} GB/s read+write")
