import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gasfm_b200 import ops
E, d = 495592, 256
x = torch.randn(E, d, device="cuda"); w = torch.randn(d, d, device="cuda") / 16; b = torch.randn(d, device="cuda")
print(os.environ.get("GASFM_GEMM_DEBUG", "0"), round(bench.timed_batches(lambda: ops.gemm_tf32x3(x, w, b), 3, 5, 3), 4))
