"""dX-style GEMM with accumulate: out += a b^T, both kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gasfm_b200 import ops  # noqa: E402

x = torch.randn(495592, 256, device="cuda")
w = torch.randn(256, 256, device="cuda") / 16
out = torch.zeros(495592, 256, device="cuda")
for name, fn in (("f16x2", ops.gemm_f16x2), ("tf32x3", ops.gemm_tf32x3)):
    plain = bench.timed_batches(lambda: fn(x, w, None, out=out), 2, 5, 3)
    acc = bench.timed_batches(lambda: fn(x, w, None, out=out, accumulate=True), 2, 5, 3)
    print(name, "store", round(plain, 4), "accumulate", round(acc, 4), flush=True)
ref = torch.randn(1000, 256, device="cuda")
o = ref.clone()
ops.gemm_f16x2(x[:1000], w, None, out=o, accumulate=True)
print("acc err", ((o.double() - (ref.double() + x[:1000].double() @ w.double().t())).abs().max()).item())
