"""A few launches of selected hot kernels at the cfg-2 shape, for `ncu --set full -k regex:<name>` (short on purpose).
Usage: python tools/kernel_probe.py [E] [dx_f16] [dx_f16_rm] [dx_tf32] [wgrad_multi] [gemm3] [peer]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import ops  # noqa: E402

which = [a for a in sys.argv[1:] if not a.isdigit()] or ["dx_f16", "wgrad_multi"]
E, d = next((int(a) for a in sys.argv[1:] if a.isdigit()), 495592), 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.relu(torch.randn(E, d, device=dev))
dys = [torch.randn(E, d, device=dev) for _ in range(3)]
w = torch.randn(d, d, device=dev) / d ** 0.5
wcat = torch.cat([w, w, w], dim=1)
amax3 = torch.stack([dy.abs().max() for dy in dys])
amax_x = x.abs().max().reshape(1)
for _ in range(3):
    if "dx_f16" in which:
        ops.gemm_f16x2_cat(dys, wcat)
    if "dx_f16_rm" in which:
        ops.gemm_f16x2_cat(dys, wcat, rowmax=[dy.abs().amax(1) for dy in dys])
    if "dx_tf32" in which:
        ops.gemm_tf32x3_cat(dys, wcat)
    if "wgrad_multi" in which:
        ops.wgrad_f16x2_multi(dys, x, amax3, amax_x)
    if "gemm3" in which:
        ops.gemm_f16x2_groups(x, [w, w, w], [None, None, None])
if "peer" in which:
    from gasfm_b200 import dist as gdist
    ranks = gdist.PeerExchange.local_group(2, dev, region_floats=1 << 20)
    acc = [torch.randn(1000, 256, device=dev) for _ in range(2)]
    mx, sm = torch.randn(1000, 4, device=dev), torch.rand(1000, 4, device=dev) + 1
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):
        for r in range(2):
            streams[r].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(streams[r]):
                ranks[r].lse_merge(acc[r], mx, sm, 4)
                ranks[r].allreduce_sum(acc[r])
torch.cuda.synchronize()
print("ok")
