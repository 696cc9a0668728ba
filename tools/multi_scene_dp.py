"""BASELINE.json configs[3]: multi-scene training step, one synthetic scene per GPU (100-500 views each),
shipped GASFM model replicated, ESFM loss, SUM gradient all-reduce over NCCL, Adam step.
    python -m torch.distributed.run --nproc-per-node N tools/multi_scene_dp.py [--steps K]
Prints one JSON line (max-over-ranks step time)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasfm_b200 import dist as gdist  # noqa: E402
from gasfm_b200.config import ConfigTree, gasfm_conf  # noqa: E402
from gasfm_b200.loss_functions import ESFMLoss  # noqa: E402
from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet  # noqa: E402
from gasfm_b200.scene import Scene  # noqa: E402
from gasfm_b200 import synthetic as gasfm_cpu  # noqa: E402  (synthetic scene generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = 100 + (400 * rank) // max(1, world - 1) if world > 1 else 300          # 100 ... 500 views
    n = 60 * m
    idx, vals = gasfm_cpu.synthetic_observations(m, n, int(0.03 * m * n), seed=100 + rank)
    scene = Scene.from_observations(idx, vals, m, n).to(dev)
    E = idx.shape[1]
    conf = gasfm_conf()
    conf["loss"] = ConfigTree.from_dict(dict(infinity_pts_margin=1e-4, hinge_loss=True, hinge_loss_weight=1,
                                             pts_grad_equalization_pre_perspective_divide=True,
                                             normalize_grad_wrt_valid_projections_only=True))
    torch.manual_seed(0)
    model = GraphAttnSfMNet(conf).to(dev)
    loss_fn = ESFMLoss(conf)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(scene), scene)
        loss.backward()                                     # batch loss = SUM over scenes (code/train.py:88)
        if world > 1:
            gdist.allreduce_gradients(model.parameters())
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        loss = step()
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / args.steps], device=dev)
    tot_e = torch.tensor([float(E)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_e)
    if rank == 0:
        print(json.dumps({"config": "cfg4 multi-scene DP", "n_gpus": world, "scenes_per_step": world,
                          "views_rank0": m, "total_observations": int(tot_e.item()), "ms_per_step": float(ms.item()),
                          "scenes_per_s": world / (float(ms.item()) / 1e3),
                          "gat_layer_edges_per_s": float(tot_e.item()) * 26 / (float(ms.item()) / 1e3),
                          "grad_allreduce_bytes": sum(p.numel() for p in model.parameters()) * 4,
                          "loss_rank0": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
