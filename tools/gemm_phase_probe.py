"""Phase isolation of the tensor-core GEMMs: GASFM_GEMM_DEBUG bits 1 = no C stores, 2 = no MMA, 4 = no B loads
(tf32x3 only), 8 = no A loads.  Usage: GASFM_GEMM_DEBUG=<bits> python tools/gemm_phase_probe.py [f16x2|tf32x3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gasfm_b200 import ops  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "f16x2"
fn = ops.gemm_f16x2 if kind == "f16x2" else ops.gemm_tf32x3
x = torch.randn(495592, 256, device="cuda")
w = torch.randn(256, 256, device="cuda") / 16
b = torch.randn(256, device="cuda")
print(kind, os.environ.get("GASFM_GEMM_DEBUG", "0"), round(bench.timed_batches(lambda: fn(x, w, b), 3, 5, 3), 4))
