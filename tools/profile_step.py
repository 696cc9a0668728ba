"""Kernel-time breakdown of one GASFM fwd+bwd step (torch.profiler, CUDA activity).
Usage: python tools/profile_step.py [cfg2|cfg3|cfg3_d256] [num_layers] > gpurun_out/step_profile.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = dict(bench.WORKLOADS[sys.argv[1]] if len(sys.argv) > 1 else bench.CFG2)
if len(sys.argv) > 2:
    cfg["num_layers"] = int(sys.argv[2])
conf, model, scene = bench.build_workload(cfg)
dev = torch.device("cuda:0")
model = model.to(dev)
scene = scene.to(dev)
for _ in range(3):
    bench.step_device(model, scene)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        bench.step_device(model, scene)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
