"""The grouped projection GEMM (3 x [E,256] = [E,256] W_g^T) alone, for timing and ncu: python tools/gemm_pair_probe.py [E] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import ops  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4987789
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(E, 256, device=dev)
ws = [torch.randn(256, 256, device=dev) / 16 for _ in range(3)]
bs = [torch.randn(256, device=dev) for _ in range(3)]
out = ops.gemm_f16x2_groups(x, ws, bs)
del out
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(reps):
    out = ops.gemm_f16x2_groups(x, ws, bs)
    del out
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
gb = E * 256 * 4 * 4 / 1e9
print(f"E={E} grouped projection GEMM: {ms:.3f} ms per launch, {gb / ms * 1e3:.0f} GB/s algorithmic")
