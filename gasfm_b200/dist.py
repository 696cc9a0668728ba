"""Track-sharded multi-GPU execution of the GASFM attention path (one process per GPU, NCCL).

The reference has no distributed code at all (single ``cuda:0``, ``code/main.py:78``); this is the
B200 design of SURVEY.md section 8(e).

Partition.  Tracks (columns of the observation matrix) are split into ``world`` contiguous blocks of
equal *edge* load, so every observation of a point lives on exactly one GPU: the column direction
(proj2scenepoint, the per-observation update, all point-level layers) needs no communication.  View
features ``[m, .]``, the global feature and all parameters are replicated.

Exchange.  A view's softmax runs over observations held by all ranks.  Each rank computes, with the
same fused kernel, the un-normalised partial ``(max_g, sum_g, acc_g)`` over its local edges; the
partials are merged like flash-attention blocks:
    M = max_g max_g          (all-reduce MAX,  m*H floats)
    [L, A] = sum_g e^{max_g-M} [sum_g, acc_g]   (all-reduce SUM, m*H*(C+1) floats)
    out = A / L + bias
The same merge (with one target) serves scenepoint2global.  Everything downstream of a merge is
replicated compute.

Backward.  The job's loss is the SUM over ranks of rank-local losses (replicated terms divided by
``world``, see ``shard_loss``).  Then every gradient in the system is a *partial* whose sum over ranks
is the true gradient: backward is linear, so replicated layers propagate partials unchanged, and the
only communication is at the merges, where a rank's local edges need the FULL output gradient --
one all-reduce SUM of ``dOut [m, H*C]`` per merge, mirroring the forward one.  A single flat
all-reduce of all parameter gradients (``allreduce_gradients``) finishes the step; that is also the
gradient exchange for scene-per-GPU data parallelism (SUM, like ``batch_loss += loss`` in
``code/train.py:88``).
"""
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from .scene import Scene
from .utils.constants import MIN_N_POINTS_PER_VIEW
from .utils.dataset_utils import AxialAggregationGraphWrapper


# ---------------------------------------------------------------------------------------------
# host logic: partition
# ---------------------------------------------------------------------------------------------
def partition_tracks(views_per_track, world):
    """Boundaries ``b[0..world]`` of contiguous track blocks with (nearly) equal edge counts:
    rank g owns tracks [b[g], b[g+1]).  Balances sum_j k_j, not n/world."""
    k = np.asarray(views_per_track, dtype=np.int64)
    n = k.size
    cum = np.concatenate(([0], np.cumsum(k)))
    total = cum[-1]
    bounds = [0]
    for g in range(1, world):
        target = total * g / world
        j = int(np.searchsorted(cum, target, side="left"))
        if j > 0 and abs(cum[j - 1] - target) <= abs(cum[min(j, n)] - target):
            j -= 1
        bounds.append(min(max(j, bounds[-1]), n))
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def shard_observations(indices, values, m, n, rank, world, bounds=None):
    """Local slice of a row-major sorted observation list: the observations of this rank's tracks,
    with local column ids.  Row-major order is preserved, so the local list is itself a valid scene."""
    indices = np.asarray(indices)
    if bounds is None:
        bounds = partition_tracks(np.bincount(indices[1], minlength=n), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    sel = (indices[1] >= lo) & (indices[1] < hi)
    local = np.stack((indices[0][sel], indices[1][sel] - lo))
    return local, np.asarray(values)[sel], lo, hi


def shard_scene(indices, values, m, n, rank, world, group=None, bounds=None):
    """Scene holding this rank's tracks.  ``view2global`` uses the GLOBAL per-view counts (views with
    >= 8 points anywhere, ``code/datasets/SceneData.py:174``); the view-aggregation and
    scenepoint2global graphs are flagged so that the model merges their partials across ``group``."""
    indices = np.asarray(indices)
    local_idx, local_vals, lo, hi = shard_observations(indices, values, m, n, rank, world, bounds)
    scene = Scene.from_observations(local_idx, local_vals, m, hi - lo)
    pts_per_view = np.bincount(indices[0], minlength=m)
    scene.x.pts_per_cam = torch.from_numpy(pts_per_view).unsqueeze(1)       # global counts (replicated)
    rows = torch.from_numpy(np.nonzero(pts_per_view >= MIN_N_POINTS_PER_VIEW)[0])
    scene.graph_wrappers["view2global"] = AxialAggregationGraphWrapper(
        m, 1, 0, valid_indices=torch.stack((rows, torch.zeros_like(rows))))
    scene.shard = ShardInfo(rank, world, lo, hi, n, group)
    scene.graph_wrappers["proj2view"].shard = scene.shard
    scene.graph_wrappers["scenepoint2global"].shard = scene.shard
    return scene


class ShardInfo:
    def __init__(self, rank, world, col_begin, col_end, n_global, group=None):
        self.rank, self.world = rank, world
        self.col_begin, self.col_end, self.n_global = col_begin, col_end, n_global
        self.group = group


# ---------------------------------------------------------------------------------------------
# collectives
# ---------------------------------------------------------------------------------------------
def lse_merge(acc, seg_max, seg_sum, heads, group=None):
    """Merge per-rank un-normalised softmax partials.  acc [T,H*C] = sum_e e^{s-max} x_e,
    seg_max / seg_sum [T,H].  Returns (normalised out [T,H*C] without bias, M [T,H], L [T,H]),
    identical on every rank.  Device-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
    T, hc = acc.shape
    M = seg_max.clone()
    dist.all_reduce(M, op=dist.ReduceOp.MAX, group=group)
    M_safe = torch.where(torch.isinf(M), torch.zeros_like(M), M)
    scale = torch.exp(seg_max - M_safe)                                   # 0 for ranks without edges
    packed = torch.cat(((seg_sum * scale).unsqueeze(-1), acc.view(T, heads, -1) * scale.unsqueeze(-1)), dim=-1)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    L = packed[..., 0].contiguous()
    inv = torch.where(L > 0, 1.0 / L.clamp_min(1e-38), torch.zeros_like(L))
    out = (packed[..., 1:] * inv.unsqueeze(-1)).reshape(T, hc)
    return out, M, L


class CudaEdgeBackend:
    """Local compute of the sharded GAT on the sm_100a kernels."""

    @staticmethod
    def partial(XL, XR, att, plan, heads):
        from . import ops
        return ops.gat_edge_partial(XL, XR, att, plan, heads)

    @staticmethod
    def backward(XL, XR, att, out_nobias, M, L, d_out, plan, heads):
        from . import ops
        return ops.gat_edge_backward_raw(XL, XR, att, out_nobias, M, L, d_out, plan, heads)


class _ShardedGat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, XL, XR, att, bias, plan, heads, group, backend):
        acc, mx, sm = backend.partial(XL, XR, att, plan, heads)
        out_nobias, M, L = lse_merge(acc, mx, sm, heads, group)
        ctx.save_for_backward(XL, XR, att, out_nobias, M, L)
        ctx.plan, ctx.heads, ctx.group, ctx.backend, ctx.has_bias = plan, heads, group, backend, bias is not None
        return out_nobias if bias is None else out_nobias + bias

    @staticmethod
    def backward(ctx, d_out):
        from . import ops
        XL, XR, att, out_nobias, M, L = ctx.saved_tensors
        d_bias = ops.col_sum(d_out) if ctx.has_bias else None               # partial (sums to the true grad)
        d_full = d_out.contiguous().clone()
        dist.all_reduce(d_full, op=dist.ReduceOp.SUM, group=ctx.group)    # local edges need the full dOut
        dXL, dXR, datt = ctx.backend.backward(XL, XR, att, out_nobias, M, L, d_full, ctx.plan, ctx.heads)
        return dXL, dXR, datt.view(att.shape), d_bias, None, None, None, None


def sharded_gat(XL, XR, att, bias, plan, heads, group=None, backend=CudaEdgeBackend):
    return _ShardedGat.apply(XL, XR, att, bias, plan, heads, group, backend)


def shard_loss(replicated_terms, local_terms, world):
    """Rank-local loss whose sum over ranks is the job's loss: terms computed identically on every
    rank (from replicated predictions such as ``Ps_norm``) are divided by ``world``."""
    return replicated_terms / world + local_terms


def allreduce_gradients(parameters, group=None):
    """One flat all-reduce SUM over all parameter gradients (parameters that received no gradient on
    a rank contribute zeros)."""
    params = [p for p in parameters if p.requires_grad]
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = flat[off:off + n].view_as(p)
        off += n
    return flat


def gather_points(pts_local, shard, group=None):
    """All-gather the per-rank ``pts3D [4, n_local]`` blocks into ``[4, n]`` (track order)."""
    sizes = [torch.zeros(1, dtype=torch.int64, device=pts_local.device) for _ in range(shard.world)]
    dist.all_gather(sizes, torch.tensor([pts_local.shape[1]], dtype=torch.int64, device=pts_local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros(pts_local.shape[0], pad, dtype=pts_local.dtype, device=pts_local.device)
    buf[:, : pts_local.shape[1]] = pts_local
    parts = [torch.empty_like(buf) for _ in range(shard.world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:, :s] for p, s in zip(parts, sizes)], dim=1)


# ---------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): one scene of N x 50k tracks, track-sharded (weak scaling)
# ---------------------------------------------------------------------------------------------
def bench_main(args, cfg, workload_config, ClockSampler, measured_peaks, timed, surrogate_loss):
    from . import _lib
    from .config import gasfm_conf
    from .models.graph_attn_sfm import GraphAttnSfMNet
    from oracle import gasfm_cpu  # synthetic scene generator only

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(cfg)
    weak = getattr(args, "scaling", "weak") == "weak"
    n_total, obs_total = (cfg["n"] * world, cfg["n_obs"] * world) if weak else (cfg["n"], cfg["n_obs"])
    idx, vals = gasfm_cpu.synthetic_observations(cfg["m"], n_total, obs_total, cfg["seed"])
    E_total = idx.shape[1]
    scene_host = shard_scene(idx, vals, cfg["m"], n_total, rank, world).pin_memory()
    torch.manual_seed(cfg["seed"])
    model = GraphAttnSfMNet(gasfm_conf(n_feat_proj=cfg["n_feat_proj"], num_layers=cfg["num_layers"])).to(dev)
    n_gat = 2 * (cfg["num_layers"] + 1)
    scene_dev = scene_host.to(dev)

    def step(scene):
        model.zero_grad(set_to_none=True)
        out = model(scene)
        loss = shard_loss(out["Ps_norm"].square().mean(), out["pts3D"].square().sum() / (4 * n_total), world)
        loss.backward()
        allreduce_gradients(model.parameters())
        return out, loss

    step(scene_dev)
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    eager_ms = timed(lambda: step(scene_dev), max(2, args.steps // 2), args.warmup, sync_dist=True)
    launches = (_lib.launch_count - l0) // (max(2, args.steps // 2) + args.warmup)
    # NOTE: the sharded step is timed eagerly.  Capturing it as a CUDA graph (gasfm_b200.graphs) deadlocked with
    # NCCL collectives inside the capture on this stack (torch 2.11 / NCCL 2.28, 2 ranks), so N>1 stays eager.
    step_fn, graphed = (lambda: step(scene_dev)), False
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(step_fn, args.steps, args.warmup, sync_dist=True)
    clocks = sampler.stop()

    holder = {}

    def step_e2e():
        out, loss = step(scene_host.to(dev, non_blocking=True))
        holder["Ps"] = out["Ps_norm"].detach().cpu()
        holder["pts"] = out["pts3D"].detach().cpu()
        holder["loss"] = float(loss.detach())
    e2e_ms = timed(step_e2e, args.steps, args.warmup, sync_dist=True)
    h2d = torch.tensor([scene_host.x.values.numel() * 4 + scene_host.x.indices.numel() * 8], device=dev, dtype=torch.float64)
    d2h = torch.tensor([holder["Ps"].numel() * 4 + holder["pts"].numel() * 4 + 4], device=dev, dtype=torch.float64)
    dist.all_reduce(h2d)
    dist.all_reduce(d2h)
    if rank == 0:
        wc = workload_config(cfg, E_total, world)
        wc["workload"] = (f"{cfg['name']}{' per GPU, weak scaling' if weak else ', strong scaling'}: ONE scene of {cfg['m']} views x {n_total} points, "
                          f"E={E_total} observations, tracks sharded over {world} GPUs (per-view softmax statistics "
                          f"merged by NCCL all-reduce), n_feat_proj={cfg['n_feat_proj']}, 4 heads, {cfg['num_layers']} layers")
        line = {"metric": "gat_layer_edges_per_sec_fwd_bwd", "value": E_total * n_gat / (ms / 1e3), "unit": "edges/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": wc,
                "e2e": {"value": E_total * n_gat / (e2e_ms / 1e3), "unit": "edges/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(d2h.item())},
                "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches), "clocks": clocks,
                "cuda_graph": graphed, "eager_ms_per_step": eager_ms, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
