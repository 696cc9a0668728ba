"""LayerNorm + projections of a block: separate LayerNorm kernel vs normalisation inside the GEMM's operand producer.
python tools/ln_gemm_probe.py [E]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gasfm_b200 import ops  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 495592
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(E, 256, device=dev)
gamma, beta = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev) * 0.1
ws = [torch.randn(256, 256, device=dev) / 16 for _ in range(3)]
bs = [torch.randn(256, device=dev) for _ in range(3)]
t = bench.timed_batches
res = {
    "ln_relu_ms": t(lambda: ops.ln_relu(x, gamma, beta, 1e-5)),
    "gemm3_ms": t(lambda: ops.gemm_f16x2_groups(x, ws, bs)),
    "gemm3_ln_ms": t(lambda: ops.gemm_f16x2_groups_ln(x, gamma, beta, 1e-5, ws, bs)),
    "gemm3_ln_y_ms": t(lambda: ops.gemm_f16x2_groups_ln(x, gamma, beta, 1e-5, ws, bs, want_y=True)),
    "gemm2_ms": t(lambda: ops.gemm_f16x2_groups(x, ws[:2], bs[:2])),
    "gemm2_ln_y_ms": t(lambda: ops.gemm_f16x2_groups_ln(x, gamma, beta, 1e-5, ws[:2], bs[:2], want_y=True)),
}
print({k: round(v, 4) for k, v in res.items()})
