"""The CTA-pair GEMMs alone, for timing and ncu:
  python tools/gemm_pair_probe.py [E] [reps] [groups|ln|cat]
groups: 3 x [E,256] = x W_g^T + b_g;  ln: the same with LayerNorm + ReLU inside the operand producer (as the model runs it);
cat: dX[E,256] = [dY0 | dY1 | dY2] Wcat^T with row maxima from upstream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import ops  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4987789
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "groups"
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(E, 256, device=dev)
ws = [torch.randn(256, 256, device=dev) / 16 for _ in range(3)]
bs = [torch.randn(256, device=dev) for _ in range(3)]
gamma, beta = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev) * 0.1
if mode == "cat":
    dys = [x, torch.randn(E, 256, device=dev), torch.randn(E, 256, device=dev)]      # three different matrices, as in backward
    rms = [d.abs().amax(dim=1) for d in dys]
    wcat = torch.cat(ws, dim=1)
    fn = lambda: ops.gemm_f16x2_cat(dys, wcat, rowmax=rms)                    # noqa: E731
elif mode == "ln":
    fn = lambda: ops.gemm_f16x2_groups_ln(x, gamma, beta, 1e-5, ws, bs)       # noqa: E731
else:
    fn = lambda: ops.gemm_f16x2_groups(x, ws, bs)                             # noqa: E731
out = fn()
del out
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(reps):
    out = fn()
    del out
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
gb = E * 256 * 4 * 4 / 1e9
print(f"E={E} {mode}: {ms:.3f} ms per launch, {gb / ms * 1e3:.0f} GB/s algorithmic")
