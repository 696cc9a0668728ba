"""The N>1 path on CPU: world_size-2 gloo processes exercise the track partition, the log-sum-exp
merge of per-rank softmax partials and the "every gradient is a partial" backward rule of
gasfm_b200.dist, with the oracle standing in for the CUDA kernels (injected backend)."""
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
        bounds = gdist.partition_tracks(np.bincount(idx[1], minlength=N_TRACKS), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        sel = torch.from_numpy((idx[1] >= lo) & (idx[1] < hi))
        xl = XL[sel].clone().requires_grad_(True)
        xr, a, b = XR.clone().requires_grad_(True), att.clone().requires_grad_(True), bias.clone().requires_grad_(True)
        plan = {"rows": torch.from_numpy(idx[0])[sel], "T": M_VIEWS}
        out = gdist.sharded_gat(xl, xr, a, b, plan, H, None, OracleEdgeBackend)
        loss = gdist.shard_loss((out * W).sum(), (xl ** 2).sum() * 0.1, world)
        loss.backward()
        flat = gdist.allreduce_gradients([xr, a, b])
        assert flat.numel() == xr.numel() + a.numel() + b.numel()
        pts = torch.arange(4 * int(hi - lo), dtype=torch.float64).reshape(4, -1) + 1000 * rank
        gathered = gdist.gather_points(pts, gdist.ShardInfo(rank, world, int(lo), int(hi), N_TRACKS))
        torch.save(dict(out=out.detach(), sel=sel, dxl=xl.grad, dxr=xr.grad, datt=a.grad, dbias=b.grad,
                        gathered=gathered), f"{result_file}.{rank}")
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
        assert torch.allclose(r["dxr"], dxr, rtol=1e-10, atol=1e-12)              # partials summed to the true gradients
        assert torch.allclose(r["datt"], datt, rtol=1e-10, atol=1e-12)
        assert torch.allclose(r["dbias"], dbias, rtol=1e-10, atol=1e-12)
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
            out, Mx, L = gdist.lse_merge(acc, mx, sm, heads=1)
            assert torch.equal(out, torch.tensor([[1.0, 2.0], [0.0, 0.0]]))     # empty segment stays 0 (bias added later)
            assert torch.equal(L, sm) and torch.equal(Mx, mx)
        finally:
            dist.destroy_process_group()
