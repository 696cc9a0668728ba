"""Accuracy and speed of the scaled 2 x FP16 GEMM against fp64 / the 3xTF32 kernel / cuBLAS fp32 (GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gasfm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [(128, 16, 32), (1000, 256, 256), (300, 64, 104), (20000, 256, 256), (70000, 64, 256), (50000, 256, 64),
          (495592, 256, 256), (495592, 32, 32)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for M, N, K in shapes:
    a = torch.randn(M, K, device=dev) * 10.0 ** torch.randint(-6, 7, (M, 1), device=dev).float()
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    ref = a.double() @ w.double().t() + b.double()
    bound = a.double().norm(dim=1, keepdim=True) * w.double().norm(dim=1).unsqueeze(0) + b.double().abs() + 1e-300
    errs = {}
    for name, fn in (("f16x2", ops.gemm_f16x2), ("tf32x3", ops.gemm_tf32x3), ("cublas", torch.nn.functional.linear)):
        c = fn(a, w, b)
        torch.cuda.synchronize()
        errs[name] = ((c.double() - ref).abs() / bound).max().item()
    line = f"M={M} N={N} K={K}: " + " ".join(f"err[{k}]={v:.2e}" for k, v in errs.items())
    if M >= 20000:
        for name, fn in (("f16x2", ops.gemm_f16x2), ("tf32x3", ops.gemm_tf32x3)):
            ms = bench.timed_batches(lambda: fn(a, w, b), 2, 5, 3)
            line += f" | {name} {ms:.3f} ms ({(M*K+M*N)*4/ms/1e6:.0f} GB/s)"
    print(line, flush=True)
