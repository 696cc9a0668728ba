"""Timeline of cluster 0 of the cta_group::2 projection GEMM: SM-clock timestamps per role and virtual tile (3 groups)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import _lib, ops  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 495592
x = torch.randn(E, 256, device="cuda")
ws = [torch.randn(256, 256, device="cuda") / 16 for _ in range(3)]
bs = [torch.randn(256, device="cuda") for _ in range(3)]
for _ in range(3):
    ops.gemm_f16x2_groups(x, ws, bs)
trace = torch.zeros(2 * 4 * 32 * 16, dtype=torch.int64, device="cuda")
_lib.call("gasfm_debug_set_gemm_trace", trace.data_ptr())
ops.gemm_f16x2_groups(x, ws, bs)
torch.cuda.synchronize()
_lib.call("gasfm_debug_set_gemm_trace", None)
t = trace.cpu().view(2, 4, 32, 16)
for cta in range(2):
    tc = t[cta]
    t0 = int(tc[tc > 0].min())
    f = lambda v: "%6d" % (int(v) - t0) if v > 0 else "     -"   # noqa: E731
    print(f"=== CTA {cta}: SM cycles since its first event")
    print("vt | P: a_empty0..3 | arrive0..3 || T: b_empty0..3 || M: tmem_empty | b_full0..3 | a_full0..3 | peer0..3 || E: start | chunks0..7 | end")
    for vt in range(int(os.environ.get("TRACE_TILES", "18"))):
        p, m, e, tm = tc[0, vt], tc[1, vt], tc[2, vt], tc[3, vt]
        print("%2d" % vt, "| P", *[f(v) for v in p[0:4]], "|", *[f(v) for v in p[4:8]], "|| T", *[f(v) for v in tm[0:4]], "|| M", f(m[0]), "|",
              *[f(v) for v in m[1:5]], "|", *[f(v) for v in m[5:9]], "|", *[f(v) for v in m[9:13]], "|| E", f(e[0]), "|", *[f(v) for v in e[2:10]], "|", f(e[1]))
