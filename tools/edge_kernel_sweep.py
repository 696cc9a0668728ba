"""BASELINE.json configs[4]: edge-attention kernel sweep -- observations 1e5..2e7, density 5-60 %, head dim 32/64
(4 heads), fp32 and bf16 storage of the projected sources -- HBM GB/s of algorithmic bytes vs the measured peak.
One JSON line per case.  Algorithmic bytes with B = bytes per stored element of XL / dXL (4 or 2):
fwd E (HC B + 4) + T (2 HC 4 + 8 H), bwd E (2 HC B + 4) + T (4 HC 4 + 8 H)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gasfm_b200 import ops  # noqa: E402
from gasfm_b200.index import ObservationIndex  # noqa: E402
from gasfm_b200 import synthetic as gasfm_cpu  # noqa: E402  (synthetic scene generator)

peaks, kind = bench.measured_peaks()
peak = float(peaks["hbm_gbs"])
dev = torch.device("cuda:0")
H = 4
sizes = [100_000, 1_000_000, 5_000_000, 20_000_000]
if len(sys.argv) > 1:
    sizes = [int(float(a)) for a in sys.argv[1:]]
for E_target in sizes:
    for rho in (0.05, 0.20, 0.60):
        m = max(16, int(round(math.sqrt(E_target / (50.0 * rho)))))
        n = 50 * m
        idx, _ = gasfm_cpu.synthetic_observations(m, n, E_target, seed=1, banded=False)
        E = idx.shape[1]
        oi = ObservationIndex(torch.from_numpy(idx).to(dev), m, n)
        for C in (32, 64):
            HC = H * C
            if E * HC * 4 * 3 > 60e9:
                continue
            XL32 = torch.randn(E, HC, device=dev)
            att = torch.randn(1, H, C, device=dev) * 0.2
            for dtype, B in (("f32", 4), ("bf16", 2)):
                XL = XL32 if B == 4 else XL32.to(torch.bfloat16)
                row = {"E": E, "m": m, "n": n, "density": round(E / (m * n), 3), "head_dim": C, "dtype": dtype}
                for name, plan, T in (("tracks", oi.by_track, n), ("views", oi.by_view, m)):
                    XR = torch.randn(T, HC, device=dev)
                    out, mx, sm = ops.gat_edge_partial(XL, XR, att, plan, H)
                    out = out / sm.repeat_interleave(C, dim=1).clamp_min(1e-30)
                    dO = torch.randn(T, HC, device=dev)
                    f = bench.timed_batches(lambda: ops.gat_edge_partial(XL, XR, att, plan, H), 3, 3, 3)
                    b = bench.timed_batches(lambda: ops.gat_edge_backward_raw(XL, XR, att, out, mx, sm, dO, plan, H), 3, 3, 3)
                    fb = E * (HC * B + 4) + T * (2 * HC * 4 + 8 * H)
                    bb = E * (2 * HC * B + 4) + T * (4 * HC * 4 + 8 * H)
                    row[f"fwd_{name}"] = [round(f, 4), round(fb / f / 1e6 / peak, 3)]
                    row[f"bwd_{name}"] = [round(b, 4), round(bb / b / 1e6 / peak, 3)]
                    del XR, out, dO
                print(json.dumps(row), flush=True)
                del XL
            del XL32
        del oi
        torch.cuda.empty_cache()
