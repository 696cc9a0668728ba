"""Edge kernels alone at the cfg-2 (or given) shapes; prints one JSON line with CUDA-event timings.
Usage: python tools/edge_kernel_bench.py [m n n_obs n_feat_proj]   (short enough to run under ncu)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = dict(bench.CFG2)
if len(sys.argv) >= 5:
    cfg.update(m=int(sys.argv[1]), n=int(sys.argv[2]), n_obs=int(sys.argv[3]), n_feat_proj=int(sys.argv[4]))
peaks, kind = bench.measured_peaks()
print(json.dumps(bench.kernel_roofline(cfg, peaks, kind)))
