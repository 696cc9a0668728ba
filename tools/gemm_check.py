"""Accuracy and speed of the tcgen05 3xTF32 GEMM against fp64 / cuBLAS fp32 (run on the GPU box)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [(128, 16, 32), (1000, 256, 256), (4097, 32, 32), (300, 64, 100), (20000, 256, 260), (495592, 256, 256), (495592, 32, 32)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for M, N, K in shapes:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    ref = (a.double() @ w.double().t() + b.double())
    t0 = time.time()
    c = ops.gemm_tf32x3(a, w, b)
    torch.cuda.synchronize()
    err = ((c.double() - ref).abs().max() / ref.abs().max()).item()
    c32 = torch.nn.functional.linear(a, w, b)
    err32 = ((c32.double() - ref).abs().max() / ref.abs().max()).item()
    torch.backends.cuda.matmul.allow_tf32 = True
    c19 = torch.nn.functional.linear(a, w, b)
    torch.backends.cuda.matmul.allow_tf32 = False
    err19 = ((c19.double() - ref).abs().max() / ref.abs().max()).item()
    def tm(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    hi, lo = ops._split_tf32(w)
    t_ours = tm(lambda: ops.gemm_tf32x3(a, w, b))
    t_cublas = tm(lambda: torch.nn.functional.linear(a, w, b))
    flops = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: err 3xTF32 {err:.2e} | cuBLAS fp32 {err32:.2e} | cuBLAS tf32 {err19:.2e} || "
          f"ours {t_ours:.3f} ms ({flops/t_ours/1e9:.1f} TFLOP/s, {(M*K+M*N)*4/t_ours/1e6:.0f} GB/s) cuBLAS fp32 {t_cublas:.3f} ms", flush=True)

print("---- weight gradient dW = dY^T X ----")
for E, Nout, Kout in [(16, 32, 32), (1000, 256, 256), (4099, 32, 32), (3000, 64, 48), (20000, 256, 256), (495592, 256, 256), (495592, 32, 32), (495592, 256, 32)]:
    dy = torch.randn(E, Nout, device=dev)
    x = torch.randn(E, Kout, device=dev)
    ref = dy.double().t() @ x.double()
    got = ops.wgrad_tf32x3(dy, x)
    torch.cuda.synchronize()
    err = ((got.double() - ref).abs().max() / ref.abs().max()).item()
    c32 = dy.t() @ x
    err32 = ((c32.double() - ref).abs().max() / ref.abs().max()).item()
    t_ours = tm(lambda: ops.wgrad_tf32x3(dy, x))
    t_cublas = tm(lambda: dy.t() @ x)
    print(f"E={E} Nout={Nout} Kout={Kout}: err 3xTF32 {err:.2e} | cuBLAS fp32 {err32:.2e} || ours {t_ours:.3f} ms "
          f"({2.0*E*Nout*Kout/t_ours/1e9:.1f} TFLOP/s, {(E*Nout+E*Kout)*4/t_ours/1e6:.0f} GB/s) cuBLAS fp32 {t_cublas:.3f} ms", flush=True)
