"""Accuracy and speed of the fp16 weight-gradient kernel against fp64 / the 3xTF32 kernel / cuBLAS fp32 (GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gasfm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [(64, 128, 64), (1000, 256, 256), (4099, 128, 192), (70001, 256, 64), (495592, 256, 256)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for E, Nout, Kout in shapes:
    dy = torch.randn(E, Nout, device=dev) * 10.0 ** torch.randint(-5, 1, (E, 1), device=dev).float() * 1e-3
    x = torch.relu(torch.randn(E, Kout, device=dev) * 3)
    ref = dy.double().t() @ x.double()
    ady, ax = dy.abs().max().reshape(1), x.abs().max().reshape(1)
    dw, db = ops.wgrad_f16x2(dy, x, ady, ax, with_bias=True)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (dw.double() - ref).abs().max().item() / scale
    err3 = (ops.wgrad_tf32x3(dy, x).double() - ref).abs().max().item() / scale
    err32 = ((dy.t() @ x).double() - ref).abs().max().item() / scale
    errb = (db.double() - dy.double().sum(0)).abs().max().item() / dy.double().sum(0).abs().max().item()
    line = f"E={E} Nout={Nout} Kout={Kout}: err f16x2 {err:.2e} | 3xTF32 {err3:.2e} | cuBLAS fp32 {err32:.2e} | db {errb:.1e}"
    if E >= 70000:
        t16 = bench.timed_batches(lambda: ops.wgrad_f16x2(dy, x, ady, ax), 2, 5, 3)
        t32 = bench.timed_batches(lambda: ops.wgrad_tf32x3(dy, x), 2, 5, 3)
        line += f" || f16x2 {t16:.3f} ms ({(E*Nout+E*Kout)*4/t16/1e6:.0f} GB/s) 3xTF32 {t32:.3f} ms"
    print(line, flush=True)
