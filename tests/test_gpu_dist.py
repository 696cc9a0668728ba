"""The track-sharded multi-GPU path on ONE GPU (the driver's test box has one):

* the peer-memory exchange kernels (``csrc/peer_comm.cu``) with two ranks that live on two streams of the same device
  (``PeerExchange.local_group``: same kernels, same flags, no IPC),
* the full sharded model -- ``_ShardedGat`` on the CUDA edge kernels (``normalize=0`` partials, merged statistics in
  backward), the per-view gradient exchanges, ``LocalGradBucket`` -- against the unsharded model: outputs and every
  parameter gradient, with rank-empty view segments in the scene,
* the same through ``CollectiveExchange`` in two gloo processes that share the GPU.
The real multi-GPU run (IPC handles, NVLink, the step replayed as a CUDA graph on every rank) is checked by
``bench.py --gpus N`` itself (``parity`` in its JSON line): two ranks capturing graphs inside ONE process is not a
configuration the product uses, and torch serialises / deadlocks it.
"""
import copy
import os
import tempfile
import threading

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import grad_errors
from gasfm_b200 import dist as gdist
from gasfm_b200.config import gasfm_conf
from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet
from gasfm_b200.scene import Scene
from gasfm_b200.synthetic import synthetic_observations

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT_TOL, GRAD_TOL = 1e-4, 2e-3


def _on_streams(fns):
    """Run fns[r]() concurrently, each in its own thread on its own CUDA stream (rank r of a local group)."""
    errs = []

    def run(fn, stream):
        try:
            with torch.cuda.stream(stream):
                fn()
        except Exception as exc:  # surfaced in the main thread
            errs.append(exc)

    streams = [torch.cuda.Stream() for _ in fns]
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    threads = [threading.Thread(target=run, args=(fn, s)) for fn, s in zip(fns, streams)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    if errs:
        raise errs[0]


@pytest.mark.parametrize("world", [2, 4])
def test_peer_exchange_kernels_match_torch(world):
    T, H, C = 37, 4, 16
    group = gdist.PeerExchange.local_group(world, DEV, region_floats=1 << 16, timeout_s=20.0)
    torch.manual_seed(world)
    acc = [torch.randn(T, H * C, device=DEV) for _ in range(world)]
    mx = [torch.randn(T, H, device=DEV) * 3 for _ in range(world)]
    sm = [torch.rand(T, H, device=DEV) + 0.5 for _ in range(world)]
    for r in range(world):                       # targets without edges on some / all ranks
        mx[r][3 * r] = float("-inf"); sm[r][3 * r] = 0; acc[r][3 * r] = 0
        mx[r][30] = float("-inf"); sm[r][30] = 0; acc[r][30] = 0
    bias = torch.randn(H * C, device=DEV)
    vec = [torch.randn(1001, device=DEV) for _ in range(world)]       # odd length: padded to float4
    res = [None] * world

    def work(r):
        def fn():
            for _ in range(3):                   # consecutive exchanges reuse the two buffer slots
                out = group[r].lse_merge(acc[r], mx[r], sm[r], H, bias)
                s = group[r].allreduce_sum(vec[r], scale=0.5)
            group[r].barrier()
            res[r] = (out, s)
        return fn

    _on_streams([work(r) for r in range(world)])
    for ex in group:
        ex.check()
    M = torch.stack(mx).max(dim=0).values
    Ms = torch.where(torch.isinf(M), torch.zeros_like(M), M)
    w = [torch.where(torch.isinf(m), torch.zeros_like(m), torch.exp(m - Ms)) for m in mx]
    L = sum(wi * si for wi, si in zip(w, sm))
    A = sum(wi.unsqueeze(-1) * a.view(T, H, C) for wi, a in zip(w, acc))
    want = torch.where(L.unsqueeze(-1) > 0, A / L.clamp_min(1e-30).unsqueeze(-1), torch.zeros_like(A)).reshape(T, H * C) + bias
    for r in range(world):
        (out, M_r, L_r), s = res[r]
        assert torch.allclose(out, want, rtol=1e-5, atol=1e-6)
        assert torch.equal(out[30], bias)                               # empty everywhere: exactly the bias
        assert torch.equal(M_r, M) and torch.allclose(L_r, L, rtol=1e-6)
        assert torch.allclose(s, 0.5 * sum(vec), rtol=1e-6, atol=1e-6)
        assert torch.equal(out, res[0][0][0]) and torch.equal(s, res[0][1])   # bit-identical on every rank


def test_peer_exchange_times_out_instead_of_hanging():
    group = gdist.PeerExchange.local_group(2, DEV, region_floats=1 << 10, timeout_s=0.2)
    x = torch.ones(8, device=DEV)
    group[0].allreduce_sum(x)                    # rank 1 never shows up
    with pytest.raises(RuntimeError, match="timed out"):
        group[0].check()


def _model_and_scene(m, n, n_obs, seed, d=64, layers=2):
    torch.manual_seed(seed)
    model = GraphAttnSfMNet(gasfm_conf(n_feat_proj=d, n_feat_scenepoint=64, n_feat_view=128, n_feat_global=256, num_layers=layers))
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "norm" in k or k.endswith(".bias"):
                p.add_(0.1 * torch.randn_like(p))
    idx, vals = synthetic_observations(m, n, n_obs, seed)
    # a few views see only tracks of the LAST quarter: their segments are empty on the other ranks
    keep = ~((idx[0] < 3) & (idx[1] < 3 * n // 4))
    idx, vals = idx[:, keep], vals[keep]
    g = torch.Generator().manual_seed(seed)
    return model, idx, vals, torch.rand(m, 3, 4, generator=g), torch.rand(4, n, generator=g)


def _single_gpu_reference(model, idx, vals, m, n, wP, wX):
    model = copy.deepcopy(model).to(DEV)
    out = model(Scene.from_observations(idx, vals, m, n).to(DEV))
    ((out["Ps_norm"] * wP.to(DEV)).sum() + (out["pts3D"] * wX.to(DEV)).sum()).backward()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    return out["Ps_norm"].detach().cpu(), out["pts3D"].detach().cpu(), grads


def _sharded_step(model, idx, vals, m, n, rank, world, exchange, wP, wX):
    scene = gdist.shard_scene(idx, vals, m, n, rank, world, exchange).to(DEV).prepare()
    lo, hi = scene.shard.col_begin, scene.shard.col_end
    bucket = gdist.LocalGradBucket(model, exchange)
    wp, wx = wP.to(DEV), wX[:, lo:hi].to(DEV)
    torch.cuda.synchronize()

    def run():                                   # no host synchronisation in here: the ranks share one GPU
        bucket.prepare()
        out = model(scene)
        ((out["Ps_norm"] * wp).sum() + (out["pts3D"] * wx).sum()).backward()
        bucket.allreduce()
        return out
    return run, (lo, hi)


@pytest.mark.parametrize("world,d", [(2, 64), (3, 256)])
def test_track_sharded_model_matches_single_gpu_peer_kernels(world, d):
    m, n = 24, 3000
    model, idx, vals, wP, wX = _model_and_scene(m, n, 30_000, seed=world, d=d)
    ps, pts, want = _single_gpu_reference(model, idx, vals, m, n, wP, wX)
    group = gdist.PeerExchange.local_group(world, DEV, region_floats=1 << 20, timeout_s=30.0)
    models = [copy.deepcopy(model).to(DEV) for _ in range(world)]
    steps, outs = [], [None] * world
    for r in range(world):
        steps.append(_sharded_step(models[r], idx, vals, m, n, r, world, group[r], wP, wX))

    def work(r):
        def fn():
            outs[r] = steps[r][0]()
        return fn

    _on_streams([work(r) for r in range(world)])
    for ex in group:
        ex.check()
    for r in range(world):
        lo, hi = steps[r][1]
        assert (outs[r]["Ps_norm"].detach().cpu() - ps).abs().max() < OUT_TOL * max(1.0, ps.abs().max())
        assert (outs[r]["pts3D"].detach().cpu() - pts[:, lo:hi]).abs().max() < OUT_TOL * max(1.0, pts.abs().max())
        assert torch.equal(outs[r]["Ps_norm"], outs[0]["Ps_norm"])       # replicated results are bit-identical
        got = {k: p.grad.detach().cpu().numpy() for k, p in models[r].named_parameters()}
        worst, key = grad_errors(got, want)
        assert worst < GRAD_TOL, (r, key, worst)
    # replicated parameters hold identical gradients on every rank without any exchange
    for (k, p0), (_, p1) in zip(models[0].named_parameters(), models[1].named_parameters()):
        if not gdist.is_local_parameter(k):
            assert torch.equal(p0.grad, p1.grad), k


def _gloo_worker(rank, world, init_file, result_file, payload):
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        model, idx, vals, wP, wX, m, n = payload
        ex = gdist.CollectiveExchange(torch.device(DEV))
        model = model.to(DEV)
        run, (lo, hi) = _sharded_step(model, idx, vals, m, n, rank, world, ex, wP, wX)
        out = run()
        torch.cuda.synchronize()
        torch.save(dict(ps=out["Ps_norm"].detach().cpu(), pts=out["pts3D"].detach().cpu(), lo=lo, hi=hi,
                        grads={k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}), f"{result_file}.{rank}")
    finally:
        dist.destroy_process_group()


def test_track_sharded_model_matches_single_gpu_collective_arm():
    world, m, n = 2, 24, 3000
    model, idx, vals, wP, wX = _model_and_scene(m, n, 30_000, seed=7, d=64)
    ps, pts, want = _single_gpu_reference(model, idx, vals, m, n, wP, wX)
    with tempfile.TemporaryDirectory() as tmp:
        init_file, result_file = os.path.join(tmp, "init"), os.path.join(tmp, "res")
        mp.spawn(_gloo_worker, args=(world, init_file, result_file, (model, idx, vals, wP, wX, m, n)), nprocs=world, join=True)
        res = [torch.load(f"{result_file}.{r}", weights_only=False) for r in range(world)]
    for r in res:
        assert (r["ps"] - ps).abs().max() < OUT_TOL * max(1.0, ps.abs().max())
        assert (r["pts"] - pts[:, r["lo"]:r["hi"]]).abs().max() < OUT_TOL * max(1.0, pts.abs().max())
        worst, key = grad_errors(r["grads"], want)
        assert worst < GRAD_TOL, (key, worst)
    assert np.array_equal(res[0]["ps"].numpy(), res[1]["ps"].numpy())
