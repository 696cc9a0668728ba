"""The N>1 path on CPU: world_size-2 gloo processes exercise the track partition, the log-sum-exp
merge of per-rank softmax partials and the "replicated tensors carry full gradients" backward rule of
gasfm_b200.dist, with the oracle standing in for the CUDA kernels (injected backend + exchange)."""
import os
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gasfm_b200 import dist as gdist
from oracle import gasfm_cpu
from oracle.gatv2conv import gatv2_edge_softmax_aggregate

H, C, M_VIEWS, N_TRACKS = 4, 4, 10, 120


class TorchExchange:
    """CPU stand-in for gasfm_b200.dist.PeerExchange (same contract: rank-ordered, identical on every rank)."""

    def __init__(self, group=None):
        self.group, self.world, self.rank = group, dist.get_world_size(group), dist.get_rank(group)

    def _gather(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        return parts

    def lse_merge(self, acc, seg_max, seg_sum, heads, bias=None):
        T, hc = acc.shape
        accs, mxs, sms = self._gather(acc), self._gather(seg_max), self._gather(seg_sum)
        M = torch.stack(mxs).max(dim=0).values
        L, A = torch.zeros_like(M), torch.zeros_like(acc).view(T, heads, -1)
        for a, mx, sm in zip(accs, mxs, sms):
            w = torch.where(torch.isinf(mx), torch.zeros_like(mx), torch.exp(mx - torch.where(torch.isinf(M), torch.zeros_like(M), M)))
            L = L + w * sm
            A = A + w.unsqueeze(-1) * a.view(T, heads, -1)
        out = torch.where(L.unsqueeze(-1) > 0, A / L.clamp_min(1e-300).unsqueeze(-1), torch.zeros_like(A)).reshape(T, hc)
        return (out if bias is None else out + bias), M, L

    def allreduce_sum(self, t, out=None, scale=1.0):
        res = sum(self._gather(t)) * scale
        if out is not None:
            out.copy_(res)
            return out
        return res


class OracleEdgeBackend:
    """CPU stand-in for the kernels: same contract as gasfm_b200.dist.CudaEdgeBackend, torch fp64."""

    @staticmethod
    def _scores(XL, XR, att, rows):
        z = XL.view(-1, H, C) + XR.view(-1, H, C)[rows]
        return (torch.nn.functional.leaky_relu(z, 0.2) * att.view(1, H, C)).sum(-1), z

    @staticmethod
    def partial(XL, XR, att, plan, heads):
        rows, T = plan["rows"], plan["T"]
        s, _ = OracleEdgeBackend._scores(XL, XR, att, rows)
        mx = torch.full((T, H), float("-inf"), dtype=XL.dtype).scatter_reduce(0, rows[:, None].expand(-1, H), s, "amax")
        p = torch.exp(s - mx[rows])
        sm = torch.zeros((T, H), dtype=XL.dtype).index_add(0, rows, p)
        acc = torch.zeros((T, H, C), dtype=XL.dtype).index_add(0, rows, p.unsqueeze(-1) * XL.view(-1, H, C))
        return acc.reshape(T, H * C), mx, sm

    @staticmethod
    def backward(XL, XR, att, out_nobias, Mx, L, d_out, plan, heads):
        rows, T = plan["rows"], plan["T"]
        s, z = OracleEdgeBackend._scores(XL, XR, att, rows)
        alpha = torch.exp(s - Mx[rows]) / L[rows]
        dO = d_out.view(T, H, C)
        D = (dO * out_nobias.view(T, H, C)).sum(-1)
        dalpha = (dO[rows] * XL.view(-1, H, C)).sum(-1)
        ds = alpha * (dalpha - D[rows])
        dz = ds.unsqueeze(-1) * att.view(1, H, C) * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.2))
        dXL = alpha.unsqueeze(-1) * dO[rows] + dz
        dXR = torch.zeros((T, H, C), dtype=XL.dtype).index_add(0, rows, dz)
        datt = (ds.unsqueeze(-1) * torch.nn.functional.leaky_relu(z, 0.2)).sum(0)
        return dXL.reshape(-1, H * C), dXR.reshape(T, H * C), datt.reshape(-1)


def _problem():
    idx, _ = gasfm_cpu.synthetic_observations(M_VIEWS, N_TRACKS, 700, seed=4)
    g = torch.Generator().manual_seed(0)
    E = idx.shape[1]
    XL = torch.randn(E, H * C, generator=g, dtype=torch.float64)
    XR = torch.randn(M_VIEWS, H * C, generator=g, dtype=torch.float64)
    att = torch.randn(1, H, C, generator=g, dtype=torch.float64)
    bias = torch.randn(H * C, generator=g, dtype=torch.float64)
    W = torch.randn(M_VIEWS, H * C, generator=g, dtype=torch.float64)
    return idx, XL, XR, att, bias, W


def _unsharded(idx, XL, XR, att, bias, W):
    E = XL.shape[0]
    xl = XL.clone().requires_grad_(True)
    xr, a, b = XR.clone().requires_grad_(True), att.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    nodes_l = torch.cat((xl, torch.zeros(M_VIEWS, H * C, dtype=torch.float64)))
    nodes_r = torch.cat((torch.zeros(E, H * C, dtype=torch.float64), xr))
    ei = torch.stack((torch.arange(E), E + torch.from_numpy(idx[0])))
    out = gatv2_edge_softmax_aggregate(nodes_l.view(-1, H, C), nodes_r.view(-1, H, C), a, ei).reshape(-1, H * C)[E:] + b
    loss = (out * W).sum() + (xl ** 2).sum() * 0.1
    loss.backward()
    return out.detach(), xl.grad, xr.grad, a.grad, b.grad


def _worker(rank, world, init_file, result_file):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        idx, XL, XR, att, bias, W = _problem()
        ex = TorchExchange()
        bounds = gdist.partition_tracks(np.bincount(idx[1], minlength=N_TRACKS), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        sel = torch.from_numpy((idx[1] >= lo) & (idx[1] < hi))
        xl = XL[sel].clone().requires_grad_(True)
        xr, a, b = XR.clone().requires_grad_(True), att.clone().requires_grad_(True), bias.clone().requires_grad_(True)
        plan = {"rows": torch.from_numpy(idx[0])[sel], "T": M_VIEWS}
        out = gdist.sharded_gat(xl, xr, a, b, plan, H, ex, OracleEdgeBackend)
        # replicated term (evaluated identically on every rank) + rank-local term
        loss = (out * W).sum() + (xl ** 2).sum() * 0.1
        loss.backward()
        datt_partial = a.grad.clone()
        datt = ex.allreduce_sum(a.grad)                      # att acts on local edges: partial -> summed
        # a replicated tensor consumed by local work: backward sums the ranks' partial gradients
        v = torch.arange(6, dtype=torch.float64).requires_grad_(True)
        (gdist.replicated_to_local(v, gdist.ShardInfo(rank, world, 0, 0, 0, ex)) * (rank + 1)).sum().backward()
        pts = torch.arange(4 * int(hi - lo), dtype=torch.float64).reshape(4, -1) + 1000 * rank
        gathered = gdist.gather_points(pts, gdist.ShardInfo(rank, world, int(lo), int(hi), N_TRACKS))
        torch.save(dict(out=out.detach(), sel=sel, dxl=xl.grad, dxr=xr.grad, datt=datt, datt_partial=datt_partial,
                        dbias=b.grad, gathered=gathered, vgrad=v.grad), f"{result_file}.{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_gat_equals_unsharded_world2():
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file, result_file = os.path.join(tmp, "init"), os.path.join(tmp, "res")
        mp.spawn(_worker, args=(world, init_file, result_file), nprocs=world, join=True)
        res = [torch.load(f"{result_file}.{r}") for r in range(world)]
    idx, XL, XR, att, bias, W = _problem()
    out, dxl, dxr, datt, dbias = _unsharded(idx, XL, XR, att, bias, W)
    for r in res:
        assert torch.allclose(r["out"], out, rtol=1e-12, atol=1e-12)               # merged output replicated on every rank
        assert torch.allclose(r["dxl"], dxl[r["sel"]], rtol=1e-10, atol=1e-12)    # local edges got their exact gradient
        assert torch.allclose(r["dxr"], dxr, rtol=1e-10, atol=1e-12)              # replicated query: FULL gradient, no extra reduce
        assert torch.allclose(r["dbias"], dbias, rtol=1e-10, atol=1e-12)          # replicated parameter: full gradient
        assert torch.allclose(r["datt"], datt, rtol=1e-10, atol=1e-12)            # local parameter: partials summed
        assert torch.equal(r["vgrad"], torch.full((6,), 3.0, dtype=torch.float64))
    assert torch.equal(res[0]["out"], res[1]["out"]) and torch.equal(res[0]["dxr"], res[1]["dxr"])   # bit-identical replicas
    assert not torch.allclose(res[0]["datt_partial"], datt)
    assert torch.equal(res[0]["gathered"], res[1]["gathered"])
    assert res[0]["gathered"].shape == (4, N_TRACKS)


def test_partition_balances_edges_not_tracks():
    rng = np.random.default_rng(0)
    k = np.concatenate((rng.integers(2, 4, size=5000), rng.integers(40, 60, size=500)))   # heavy tail at the end
    for world in (2, 4, 8):
        b = gdist.partition_tracks(k, world)
        assert b[0] == 0 and b[-1] == k.size and np.all(np.diff(b) >= 0)
        loads = np.array([k[b[g]:b[g + 1]].sum() for g in range(world)])
        assert loads.max() <= 1.05 * k.sum() / world
        assert np.diff(b).max() > 2 * np.diff(b).min()          # equal load means unequal track counts here


def test_shard_observations_is_a_partition_and_stays_row_major():
    idx, vals = gasfm_cpu.synthetic_observations(12, 300, 2000, seed=1)
    seen = np.zeros(idx.shape[1], dtype=bool)
    key = idx[0] * 300 + idx[1]
    for rank in range(4):
        li, lv, lo, hi = gdist.shard_observations(idx, vals, 12, 300, rank, 4)
        assert np.all(np.diff(li[0] * (hi - lo) + li[1]) > 0)   # still row-major sorted, no duplicates
        gkey = li[0] * 300 + li[1] + lo
        pos = np.searchsorted(key, gkey)
        assert np.array_equal(key[pos], gkey) and not seen[pos].any()
        assert np.array_equal(vals[pos], lv)
        seen[pos] = True
    assert seen.all()
    s = gdist.shard_scene(idx, vals, 12, 300, 1, 4)
    assert s.x.shape[0] == 12 and s.graph_wrappers["proj2view"].shard.world == 4
    assert s.x.pts_per_cam.sum().item() == idx.shape[1]        # per-view counts stay global


def test_lse_merge_single_rank_identity():
    with tempfile.TemporaryDirectory() as tmp:
        dist.init_process_group("gloo", init_method=f"file://{os.path.join(tmp, 'i')}", rank=0, world_size=1)
        try:
            acc = torch.tensor([[2.0, 4.0], [0.0, 0.0]])
            mx = torch.tensor([[0.5], [float("-inf")]])
            sm = torch.tensor([[2.0], [0.0]])
            out, Mx, L = TorchExchange().lse_merge(acc, mx, sm, heads=1)
            assert torch.equal(out, torch.tensor([[1.0, 2.0], [0.0, 0.0]]))     # empty segment stays 0 (bias added later)
            assert torch.equal(L, sm) and torch.equal(Mx, mx)
        finally:
            dist.destroy_process_group()


def test_parameter_classification_covers_the_shipped_model():
    """Every parameter of the shipped configuration is either local (applied to observation / point tensors: partial
    gradient per rank) or replicated; the local ones are the small minority that needs the gradient exchange."""
    from gasfm_b200.config import gasfm_conf
    from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet

    with torch.device("meta"):
        model = GraphAttnSfMNet(gasfm_conf(n_feat_proj=256, num_layers=3))
    names = [k for k, _ in model.named_parameters()]
    local = [k for k in names if gdist.is_local_parameter(k)]
    sizes = dict((k, p.numel()) for k, p in model.named_parameters())
    assert sum(sizes[k] for k in local) < 0.1 * sum(sizes.values())
    must_be_local = ["embed.post_embed_lin.weight", "equivariant_blocks.1.prev_projfeat_norm_layer.weight",
                     "equivariant_blocks.1.global_feature_update.proj2view.graph_conv.lin_l.weight",
                     "equivariant_blocks.1.global_feature_update.proj2view.graph_conv.att",
                     "equivariant_blocks.1.global_feature_update.proj2scenepoint.graph_conv.lin_r.weight",
                     "equivariant_blocks.1.global_feature_update.proj2scenepoint.mlp.0.weight",
                     "equivariant_blocks.1.global_feature_update.view_and_scenepoint2global.graph_conv_scenepoint2global.lin_l.weight",
                     "equivariant_blocks.1.projection_feature_update.lin_proj.weight",
                     "equivariant_blocks.1.projection_feature_update.lin_scenepoint.weight",
                     "equivariant_blocks.0.skip_projection.lin_proj.weight",
                     "final_global_update.proj2scenepoint.graph_conv.att", "scenepoint_head.4.weight"]
    must_be_replicated = ["equivariant_blocks.1.global_feature_update.proj2view.graph_conv.lin_r.weight",
                          "equivariant_blocks.1.global_feature_update.proj2view.graph_conv.bias",
                          "equivariant_blocks.1.global_feature_update.proj2view.norm_and_proj_view2proj.2.weight",
                          "equivariant_blocks.1.global_feature_update.proj2view.mlp.0.weight",
                          "equivariant_blocks.1.global_feature_update.view_and_scenepoint2global.graph_conv_scenepoint2global.lin_r.weight",
                          "equivariant_blocks.1.global_feature_update.view_and_scenepoint2global.graph_conv_scenepoint2global.bias",
                          "equivariant_blocks.1.global_feature_update.view_and_scenepoint2global.graph_conv_view2global.lin_l.weight",
                          "equivariant_blocks.1.projection_feature_update.lin_view.weight",
                          "equivariant_blocks.1.projection_feature_update.lin_global.weight",
                          "equivariant_blocks.1.projection_feature_update.view_norm_layer.weight", "view_head.0.weight"]
    for k in must_be_local:
        assert k in names and gdist.is_local_parameter(k), k
    for k in must_be_replicated:
        assert k in names and not gdist.is_local_parameter(k), k


def test_local_grad_bucket_points_gradients_at_one_flat_buffer():
    class _Ex:
        region_floats = 1 << 20

        def allreduce_sum(self, t, out=None, scale=1.0):
            out.copy_(2 * t)
            return out

    model = torch.nn.ModuleDict({"scenepoint_head": torch.nn.Linear(3, 2), "view_head": torch.nn.Linear(3, 2)})
    bucket = gdist.LocalGradBucket(model, _Ex())
    assert len(bucket.local) == 2 and len(bucket.replicated) == 2 and bucket.flat.numel() == 8
    bucket.prepare()
    x = torch.ones(4, 3)
    (model["scenepoint_head"](x).sum() + model["view_head"](x).sum()).backward()
    assert model["scenepoint_head"].weight.grad.data_ptr() == bucket.flat.data_ptr()      # accumulated in place
    bucket.allreduce()
    assert torch.equal(model["scenepoint_head"].weight.grad, torch.full((2, 3), 8.0))
    assert torch.equal(model["view_head"].weight.grad, torch.full((2, 3), 4.0))


def _reducer_worker(rank, world, init_file, result_file):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
        unused = torch.nn.Linear(2, 2)                       # never used in forward: its bucket is flushed by finish()
        model.add_module("unused", unused)
        reducer = gdist.BucketedGradReducer(model, bucket_bytes=64)     # tiny buckets: several launches per step
        assert len(reducer.buckets) > 2
        xs = [torch.full((4, 6), float(rank + 1)), torch.full((4, 6), float(rank + 3))]      # two scenes per rank
        for _ in range(2):                                   # second step: buffers are re-zeroed, hooks fire again
            reducer.prepare()
            with reducer.accumulate_only():
                model[:4](xs[0]).square().sum().backward()
            model[:4](xs[1]).square().sum().backward()
            reducer.finish()
        torch.save({k: p.grad.clone() for k, p in model.named_parameters()}, f"{result_file}.{rank}")
    finally:
        dist.destroy_process_group()


def test_bucketed_grad_reducer_sums_over_ranks_and_scenes():
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file, result_file = os.path.join(tmp, "init"), os.path.join(tmp, "res")
        mp.spawn(_reducer_worker, args=(world, init_file, result_file), nprocs=world, join=True)
        res = [torch.load(f"{result_file}.{r}") for r in range(world)]
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    for v in (1.0, 3.0, 2.0, 4.0):                           # all four scenes of the batch, summed
        model(torch.full((4, 6), v)).square().sum().backward()
    for k, p in model.named_parameters():
        for r in res:
            assert torch.allclose(r[k], p.grad, rtol=1e-5, atol=1e-6), k
    assert torch.equal(res[0]["unused.weight"], torch.zeros(2, 2))


def test_shard_scene_rejects_empty_shards():
    import pytest
    idx = np.array([[0, 0, 1, 1, 2, 2], [0, 1, 0, 2, 1, 2]], dtype=np.int64)        # 3 tracks cannot feed 8 ranks
    vals = np.zeros((6, 2), dtype=np.float32)
    with pytest.raises(ValueError, match="no observations"):
        for rank in range(8):
            gdist.shard_scene(idx, vals, 3, 3, rank, 8)
