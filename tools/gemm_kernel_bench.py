"""The tcgen05 kernels alone at the cfg-2 projection shape (E x 256 x 256); short enough for ncu.
Prints ms per launch of: the default projection GEMM (scaled 2 x FP16), the grouped (3 projections) form,
the 3xTF32 GEMM, the concatenated-dY input gradient and the weight gradient (3xTF32 and fp16)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gasfm_b200 import ops  # noqa: E402

E, d = 495592, 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(E, d, device=dev)
dy = torch.randn(E, d, device=dev)
w = torch.randn(d, d, device=dev) / d ** 0.5
b = torch.randn(d, device=dev)
amax_dy, amax_x = dy.abs().max().reshape(1), x.abs().max().reshape(1)
reps = (1, 2, 1) if len(sys.argv) > 1 and sys.argv[1] == "short" else (2, 3, 2)
res = {"gemm_f16x2_ms": bench.timed_batches(lambda: ops.gemm_f16x2(x, w, b), *reps),
       "gemm_f16x2_3groups_ms": bench.timed_batches(lambda: ops.gemm_f16x2_groups(x, [w, w, w], [b, b, b]), *reps),
       "gemm_tf32x3_ms": bench.timed_batches(lambda: ops.gemm_tf32x3(x, w, b), *reps),
       "dx_cat3_tf32x3_ms": bench.timed_batches(lambda: ops.gemm_tf32x3_cat([dy, x, dy], torch.cat([w, w, w], dim=1)), *reps),
       "wgrad_tf32x3_ms": bench.timed_batches(lambda: ops.wgrad_tf32x3(dy, x), *reps),
       "wgrad_f16x2_ms": bench.timed_batches(lambda: ops.wgrad_f16x2(dy, x, amax_dy, amax_x), *reps)}
print(json.dumps({k: round(v, 4) for k, v in res.items()}))
