"""One warm-up step and one measured eager step of the cfg2 workload (for `ncu --metrics gpu__time_duration.sum`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = dict(bench.CFG2)
conf, model, scene = bench.build_workload(cfg)
dev = torch.device("cuda:0")
model, scene = model.to(dev), scene.to(dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    bench.step_device(model, scene)
torch.cuda.synchronize()
print("ok")
