"""The two tcgen05 kernels alone at the cfg-2 projection shape (E x 256 x 256); short enough for ncu."""
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
res = {"gemm_ms": bench.timed_batches(lambda: ops.gemm_tf32x3(x, w, b), 2, 3, 2),
       "wgrad_ms": bench.timed_batches(lambda: ops.wgrad_tf32x3(dy, x), 2, 3, 2)}
print(json.dumps(res))
