"""torchrun --nproc-per-node N tools/check_sharded_model.py
Track-sharded GASFM (N GPUs, NCCL) against the same model on one GPU: predictions and the
all-reduced parameter gradients must agree to fp32 summation-order noise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import dist as gdist  # noqa: E402
from gasfm_b200.config import gasfm_conf  # noqa: E402
from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet  # noqa: E402
from gasfm_b200.scene import Scene  # noqa: E402
from oracle import gasfm_cpu  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    m, n = 40, 6000
    idx, vals = gasfm_cpu.synthetic_observations(m, n, 60000, seed=7)
    torch.manual_seed(0)
    model = GraphAttnSfMNet(gasfm_conf(n_feat_view=256, n_feat_global=512, num_layers=3)).to(dev)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "norm" in k or k.endswith(".bias"):
                p.add_(0.1 * torch.randn_like(p))
    wP = torch.linspace(0.5, 1.5, m * 12, device=dev).reshape(m, 3, 4)
    wX = torch.linspace(-1.0, 1.0, 4 * n, device=dev).reshape(4, n)

    full = Scene.from_observations(idx, vals, m, n).to(dev)
    out = model(full)
    ((out["Ps_norm"] * wP).sum() + (out["pts3D"] * wX).sum()).backward()
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    ref_P, ref_X = out["Ps_norm"].detach(), out["pts3D"].detach()

    model.zero_grad(set_to_none=True)
    sh = gdist.shard_scene(idx, vals, m, n, rank, world).to(dev)
    o2 = model(sh)
    lo, hi = sh.shard.col_begin, sh.shard.col_end
    loss = gdist.shard_loss((o2["Ps_norm"] * wP).sum(), (o2["pts3D"] * wX[:, lo:hi]).sum(), world)
    loss.backward()
    gdist.allreduce_gradients(model.parameters())
    pts = gdist.gather_points(o2["pts3D"].detach(), sh.shard)
    eP = (o2["Ps_norm"].detach() - ref_P).abs().max().item() / max(1.0, ref_P.abs().max().item())
    eX = (pts - ref_X).abs().max().item() / max(1.0, ref_X.abs().max().item())
    gscale = 1e-3 * max(float(g.abs().max()) for g in ref_grads.values())
    worst, wk = 0.0, None
    for k, p in model.named_parameters():
        e = float((p.grad - ref_grads[k]).abs().max()) / max(gscale, float(ref_grads[k].abs().max()))
        if e > worst:
            worst, wk = e, k
    print(f"[rank {rank}/{world}] E_local={sh.x.indices.shape[1]} of {idx.shape[1]}  Ps err {eP:.2e}  pts err {eX:.2e}  "
          f"worst grad err {worst:.2e} ({wk})", flush=True)
    ok = eP < 1e-4 and eX < 1e-4 and worst < 2e-3
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
