"""Timeline of cluster 0 of the cta_group::2 concatenated-dY input-gradient GEMM (3 segments, K = 768)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import _lib, ops  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 495592
dys = [torch.randn(E, 256, device="cuda") for _ in range(3)]
rms = [d.abs().amax(dim=1) for d in dys]
w = torch.randn(256, 768, device="cuda") / 16
for _ in range(3):
    ops.gemm_f16x2_cat(dys, w, None, False, rms)
trace = torch.zeros(2 * 4 * 32 * 16, dtype=torch.int64, device="cuda")
_lib.call("gasfm_debug_set_gemm_trace", trace.data_ptr())
ops.gemm_f16x2_cat(dys, w, None, False, rms)
torch.cuda.synchronize()
_lib.call("gasfm_debug_set_gemm_trace", None)
t = trace.cpu().view(2, 4, 32, 16)
for cta in range(2):
    tc = t[cta]
    t0 = int(tc[tc > 0].min())
    f = lambda v: "%6d" % (int(v) - t0) if v > 0 else "     -"   # noqa: E731
    print(f"=== CTA {cta}: SM cycles since its first event")
    print("tile | P: stage free for kb 0..11 || T: b_empty kb 0..11 || M: tmem_empty | operands ready kb 0..11 || E: start | chunks 0..7 | end")
    for it in range(int(os.environ.get("TRACE_TILES", "8"))):
        p, m, e, tm = tc[0, it], tc[1, it], tc[2, it], tc[3, it]
        print("%2d" % it, "| P", *[f(v) for v in p[0:12]], "|| T", *[f(v) for v in tm[0:12]], "|| M", f(m[0]), "|", *[f(v) for v in m[1:13]],
              "|| E", f(e[0]), "|", *[f(v) for v in e[2:10]], "|", f(e[1]))
