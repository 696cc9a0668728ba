"""Timeline of CTA 0 of the fp16x2 GEMM: SM-clock timestamps of the producer / MMA / epilogue milestones per tile."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import _lib, ops  # noqa: E402

x = torch.randn(495592, 256, device="cuda")
w = torch.randn(256, 256, device="cuda") / 16
b = torch.randn(256, device="cuda")
for _ in range(3):
    ops.gemm_f16x2(x, w, b)
trace = torch.zeros(3 * 16 * 16, dtype=torch.int64, device="cuda")
_lib.call("gasfm_debug_set_gemm_trace", trace.data_ptr())
ops.gemm_f16x2(x, w, b)
torch.cuda.synchronize()
_lib.call("gasfm_debug_set_gemm_trace", None)
t = trace.cpu().view(3, 16, 16)
t0 = int(t[t > 0].min())
us = lambda v: "%7.2f" % ((int(v) - t0) / 1965.0) if v > 0 else "      -"   # noqa: E731
print("times in us at 1965 MHz since the first event")
print("tile | P: amax | empty0..3 | arrive0..3 || M: tmem_empty | B full0..3 | A ready0..3 | commit0..3 || E: start end")
for it in range(int(os.environ.get("TRACE_TILES", "12"))):
    p, m, e = t[0, it], t[1, it], t[2, it]
    print(it, "| P", us(p[0]), "|", *[us(v) for v in p[1:5]], "|", *[us(v) for v in p[5:9]], "|| M", us(m[0]), "|",
          *[us(v) for v in m[1:5]], "|", *[us(v) for v in m[5:9]], "|", *[us(v) for v in m[9:13]], "|| E", us(e[0]), us(e[1]))
