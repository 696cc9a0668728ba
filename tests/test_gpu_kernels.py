"""Kernel-level parity on the B200: every C-ABI entry point against the CPU oracle on the same
seeded inputs.  Integer / index outputs must be bit-exact; fp32 outputs must agree with the fp64
oracle to the tolerance written next to each check."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden
from gasfm_b200 import _lib, ops
from gasfm_b200.index import ObservationIndex, SegmentPlan, plan_from_targets, single_segment_chunk
from gasfm_b200.scene import Scene
from gasfm_b200.utils import dataset_utils, sparse_utils
from oracle import gasfm_cpu, gat_edge_c

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

FP32_TOL = 2e-5      # max abs err / max(1, |ref|) for forward outputs (fp32 kernels vs fp64 oracle)
GRAD_TOL = 1e-4      # same for gradients (longer accumulation chains)


def rel_err(got, want):
    got = got.detach().double().cpu().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(1.0, np.abs(want).max())) if want.size else 0.0


# ---------------------------------------------------------------------------------------------
# index build
# ---------------------------------------------------------------------------------------------
def test_m2sparse_matches_reference_golden_bit_exact():
    g = load_golden("index_build")
    x = dataset_utils.M2sparse(torch.from_numpy(g["M"]).to(DEV), normalize=True, Ns=torch.from_numpy(g["Ns"]).to(DEV))
    assert np.array_equal(x.indices.cpu().numpy(), g["indices"])
    assert np.array_equal(x.cam_per_pts.cpu().numpy(), g["cam_per_pts"])
    assert np.array_equal(x.pts_per_cam.cpu().numpy(), g["pts_per_cam"])
    assert tuple(x.shape) == tuple(g["shape"])
    np.testing.assert_allclose(x.values.cpu().numpy(), g["values"], rtol=2e-6, atol=1e-7)
    valid = dataset_utils.get_M_valid_points(torch.from_numpy(g["M"]).to(DEV)).cpu().numpy()
    assert np.array_equal(valid, gasfm_cpu.valid_observation_mask(g["M"]))
    # host tensors are accepted too (moved to the GPU for the kernels, result returned on the host)
    xh = dataset_utils.M2sparse(torch.from_numpy(g["M"]), normalize=False)
    assert not xh.values.is_cuda and np.array_equal(xh.indices.numpy(), g["indices"])
    graphs = dataset_utils.create_axial_aggregation_graphs(x)
    for name, w in graphs.items():
        assert np.array_equal(w.edge_index.cpu().numpy(), g[f"{name}.edge_index"]), name
        assert np.array_equal(w.valid_indices.cpu().numpy(), g[f"{name}.valid_indices"]), name


@pytest.mark.parametrize("m,n,E,seed", [(20, 2000, 12000, 0), (7, 40, 150, 1), (300, 50000, 500000, 2), (64, 100, 3800, 3)])
def test_csr_csc_bit_exact(m, n, E, seed):
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, E, seed)
    idx = torch.from_numpy(idx_np).to(DEV)
    oi = ObservationIndex(idx, m, n)
    rows, cols = torch.from_numpy(idx_np[0]), torch.from_numpy(idx_np[1])
    row_ptr = torch.zeros(m + 1, dtype=torch.int64); row_ptr[1:] = torch.bincount(rows, minlength=m).cumsum(0)
    col_ptr = torch.zeros(n + 1, dtype=torch.int64); col_ptr[1:] = torch.bincount(cols, minlength=n).cumsum(0)
    perm = torch.argsort(cols, stable=True)
    assert torch.equal(oi.row_idx.cpu().long(), rows) and torch.equal(oi.col_idx.cpu().long(), cols)
    assert torch.equal(oi.row_ptr.cpu().long(), row_ptr)
    assert torch.equal(oi.col_ptr.cpu().long(), col_ptr)
    assert torch.equal(oi.csc_perm.cpu().long(), perm)
    # chunk tables of the view plan
    p = oi.by_view
    lens = (row_ptr[1:] - row_ptr[:-1])
    nchunks = (lens + p.chunk - 1) // p.chunk
    cp = torch.zeros(m + 1, dtype=torch.int64); cp[1:] = nchunks.cumsum(0)
    assert torch.equal(p.chunk_ptr.cpu().long(), cp)
    total = int(cp[-1])
    assert total <= p.max_chunks
    assert torch.equal(p.chunk_seg.cpu().long()[:total], torch.repeat_interleave(torch.arange(m), nchunks))


def test_csr_build_rejects_bad_indices():
    bad = torch.tensor([[0, 0, 1], [1, 1, 0]], dtype=torch.int64, device=DEV)     # duplicate entry
    with pytest.raises(ValueError, match="sorted"):
        ObservationIndex(bad, 2, 2)
    oob = torch.tensor([[0, 5], [0, 1]], dtype=torch.int64, device=DEV)
    with pytest.raises(ValueError, match="range"):
        ObservationIndex(oob, 2, 2)


def test_csr_build_host_entry_point():
    idx_np, _ = gasfm_cpu.synthetic_observations(9, 50, 200, 5)
    E = idx_np.shape[1]
    idx = np.ascontiguousarray(idx_np, np.int64)
    rp, cp, perm = np.zeros(10, np.int32), np.zeros(51, np.int32), np.zeros(E, np.int32)
    lib = _lib.load()
    rc = lib.gasfm_csr_build_host(idx.ctypes.data, E, 9, 50, rp.ctypes.data, cp.ctypes.data, perm.ctypes.data)
    assert rc == 0, _lib.last_error()
    assert np.array_equal(perm, np.argsort(idx[1], kind="stable"))
    assert np.array_equal(np.diff(rp), np.bincount(idx[0], minlength=9))
    assert np.array_equal(np.diff(cp), np.bincount(idx[1], minlength=50))


# ---------------------------------------------------------------------------------------------
# fused GATv2 edge attention
# ---------------------------------------------------------------------------------------------
def _gat_case(E, T, H, C, seed, empty_tail=0, bcast=False):
    rng = np.random.default_rng(seed)
    target = np.sort(rng.integers(0, max(1, T - empty_tail), size=E))
    XL = rng.standard_normal((E, H * C)).astype(np.float32)
    XR = rng.standard_normal((1 if bcast else T, H * C)).astype(np.float32)
    att = (rng.standard_normal(H * C) * 0.5).astype(np.float32)
    bias = rng.standard_normal(H * C).astype(np.float32)
    dOut = rng.standard_normal((T, H * C)).astype(np.float32)
    return target, XL, XR, att, bias, dOut


def _run_gat(plan, XL, XR, att, bias, dOut, H, C):
    t = lambda a: torch.from_numpy(a).to(DEV).requires_grad_(True)  # noqa: E731
    xl, xr, a, b = t(XL), t(XR), t(att.reshape(1, H, C)), t(bias)
    out = ops.gat_edge_attention(xl, xr, a, b, plan, H)
    out.backward(torch.from_numpy(dOut).to(DEV))
    torch.cuda.synchronize()
    return out, xl.grad, xr.grad, a.grad.reshape(-1), b.grad


def _check_gat(target, XL, XR, att, bias, dOut, T, H, C, plan, perm_rows=None):
    """perm_rows: edge e of the oracle corresponds to row perm_rows[e] of XL on the device."""
    XLo = XL if perm_rows is None else XL[perm_rows]
    ref_out, _, _ = gat_edge_c.gat_edge_fwd(XLo, XR, att, bias, target, T, H, C)
    dXL, dXR, datt, dbias = gat_edge_c.gat_edge_bwd(XLo, XR, att, target, dOut, T, H, C)
    out, gxl, gxr, gatt, gb = _run_gat(plan, XL, XR, att, bias, dOut, H, C)
    assert rel_err(out, ref_out) < FP32_TOL
    g = gxl.cpu().numpy()
    if perm_rows is not None:
        full = np.zeros_like(g, dtype=np.float64); full[perm_rows] = dXL; dXL = full
    assert rel_err(g, dXL) < GRAD_TOL
    assert rel_err(gxr, dXR) < GRAD_TOL
    assert rel_err(gatt, datt) < GRAD_TOL * 5
    assert rel_err(gb, dbias) < GRAD_TOL
    return out


HEAD_SHAPES = [(4, 8), (4, 64), (4, 1), (4, 2), (4, 4), (4, 16), (4, 32), (4, 128), (4, 256), (2, 6), (1, 5), (3, 8), (8, 4)]


@pytest.mark.parametrize("H,C", HEAD_SHAPES)
def test_gat_short_segments_through_csc_perm(H, C):
    """tracks: many short segments, rows gathered through a permutation, some targets empty"""
    m, n, E0 = 16, 300, 1500
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, E0, seed=H * 7 + C)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m, n + 5)      # 5 trailing tracks without observations
    _, XL, XR, att, bias, dOut = _gat_case(E, n + 5, H, C, seed=C)
    out = _check_gat(idx_np[1], XL, XR, att, bias, dOut, n + 5, H, C, oi.by_track)
    assert torch.equal(out[-5:].cpu(), torch.from_numpy(bias).expand(5, -1))   # empty segment -> exactly bias


@pytest.mark.parametrize("H,C", HEAD_SHAPES)
def test_gat_long_segments_chunked(H, C):
    """views: few long contiguous segments -> chunked schedule + log-sum-exp merge"""
    m, n, E0 = 9, 400, 2600
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, E0, seed=H + C)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m + 2, n)      # 2 trailing views without observations
    assert oi.by_view.chunk > 0
    _, XL, XR, att, bias, dOut = _gat_case(E, m + 2, H, C, seed=C + 1)
    _check_gat(idx_np[0], XL, XR, att, bias, dOut, m + 2, H, C, oi.by_view)


@pytest.mark.parametrize("C", [32, 64])
@pytest.mark.parametrize("direction", ["tracks", "views"])
def test_gat_bf16_stored_sources(C, direction):
    """BASELINE.json configs[4], "fp32 vs bf16": XL stored as bf16 (half the bytes per edge), arithmetic in fp32.  The kernel
    must be exact for the bf16-ROUNDED inputs (fp32 tolerances against the fp64 oracle fed with the rounded values); dXL comes
    back in bf16, i.e. within its storage rounding (2^-9 relative per element).  Both schedules, empty segments included."""
    H, m, n = 4, 12, 300
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, 2400, seed=C)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m + 1, n + 3)
    plan, T, target = (oi.by_track, n + 3, idx_np[1]) if direction == "tracks" else (oi.by_view, m + 1, idx_np[0])
    _, XL, XR, att, bias, dOut = _gat_case(E, T, H, C, seed=C + 5)
    xl16 = torch.from_numpy(XL).to(DEV).to(torch.bfloat16)
    XLr = xl16.float().cpu().numpy()                              # what the kernel actually reads
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    acc, mx, sm = ops.gat_edge_partial(xl16, t(XR), t(att.reshape(1, H, C)), plan, H)
    L = sm.repeat_interleave(C, dim=1)
    out = torch.where(L > 0, acc / L.clamp_min(1e-30), torch.zeros_like(acc))
    ref_out, ref_max, ref_sum = gat_edge_c.gat_edge_fwd(XLr, XR, att, None, target, T, H, C)
    assert rel_err(out, ref_out) < FP32_TOL
    dXL, dXR, datt, _ = gat_edge_c.gat_edge_bwd(XLr, XR, att, target, dOut, T, H, C)
    gxl, gxr, gatt = ops.gat_edge_backward_raw(xl16, t(XR), t(att.reshape(1, H, C)), out, mx, sm, t(dOut), plan, H)
    assert gxl.dtype == torch.bfloat16
    assert np.abs(gxl.float().cpu().numpy() - dXL).max() < 2.0 ** -8 * max(1.0, np.abs(dXL).max())
    assert rel_err(gxr, dXR) < GRAD_TOL
    assert rel_err(gatt, datt) < GRAD_TOL * 5
    # against the UNROUNDED fp32 inputs the storage rounding is what is left: stated, looser tolerance
    full_out, _, _ = gat_edge_c.gat_edge_fwd(XL, XR, att, None, target, T, H, C)
    assert rel_err(out, full_out) < 2e-2


@pytest.mark.parametrize("H,C,E,chunk", [(4, 16, 5000, None), (4, 256, 300, None), (4, 8, 70, 8), (4, 64, 1, 8), (2, 6, 100, 8)])
def test_gat_single_target_global_graph(H, C, E, chunk):
    """view2global / scenepoint2global: ONE segment holding every edge; with a subset permutation"""
    rng = np.random.default_rng(E)
    n_rows = E + 7
    keep = np.sort(rng.choice(n_rows, size=E, replace=False))
    seg_ptr = torch.tensor([0, E], dtype=torch.int32, device=DEV)
    plan = SegmentPlan(seg_ptr, torch.from_numpy(keep.astype(np.int32)).to(DEV), 1, E, chunk or single_segment_chunk(E), DEV)
    _, XL, XR, att, bias, dOut = _gat_case(n_rows, 1, H, C, seed=3)
    _check_gat(np.zeros(E, np.int64), XL, XR, att, bias, dOut, 1, H, C, plan, perm_rows=keep)


def test_gat_stateless_broadcast_query_and_strided_rows():
    """first block: query = lin_r.bias broadcast to every target; XL given as a strided slice"""
    H, C, m, n = 4, 8, 12, 200
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, 900, seed=9)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m, n)
    for plan, tgt, T in ((oi.by_track, idx_np[1], n), (oi.by_view, idx_np[0], m)):
        _, XL, XR, att, bias, dOut = _gat_case(E, T, H, C, seed=4, bcast=True)
        _check_gat(tgt, XL, XR, att, bias, dOut, T, H, C, plan)
        wide = torch.randn(E, 3 * H * C, device=DEV)
        sl = wide[:, H * C: 2 * H * C]
        o1 = ops.gat_edge_attention(sl, torch.from_numpy(XR).to(DEV), torch.from_numpy(att).to(DEV).view(1, H, C), None, plan, H)
        o2 = ops.gat_edge_attention(sl.contiguous(), torch.from_numpy(XR).to(DEV), torch.from_numpy(att).to(DEV).view(1, H, C), None, plan, H)
        assert torch.equal(o1, o2)


def test_gat_large_scores_are_stable():
    """softmax must survive scores of +-80 (max subtraction), like the reference's"""
    H, C, E, T = 4, 8, 4000, 3
    target, XL, XR, att, bias, dOut = _gat_case(E, T, H, C, seed=11)
    XL *= 30.0
    plan = plan_from_targets(torch.from_numpy(target).to(DEV), T)
    out = _check_gat(target, XL, XR, att, bias, dOut, T, H, C, plan)
    assert torch.isfinite(out).all()


def test_gat_partial_statistics_merge_like_flash_attention():
    """normalize=0 partials of two edge shards merge to the full result (the multi-GPU combine)"""
    H, C, m, n = 4, 16, 10, 240
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, 1500, seed=21)
    E = idx_np.shape[1]
    _, XL, XR, att, bias, _ = _gat_case(E, m, H, C, seed=5)
    full, _, _ = gat_edge_c.gat_edge_fwd(XL, XR, att, None, idx_np[0], m, H, C)
    half = idx_np[1] < n // 2
    parts = []
    for sel in (half, ~half):
        sub = idx_np[:, sel]
        oi = ObservationIndex(torch.from_numpy(np.ascontiguousarray(sub)).to(DEV), m, n)
        o, mx, sm = ops.gat_edge_partial(torch.from_numpy(XL[sel]).to(DEV), torch.from_numpy(XR).to(DEV),
                                         torch.from_numpy(att).to(DEV), oi.by_view, H)
        parts.append((o.view(m, H, C), mx, sm))
    M = torch.maximum(parts[0][1], parts[1][1])
    num = sum(o * torch.exp(mx - M).unsqueeze(-1) for o, mx, _ in parts)
    den = sum(sm * torch.exp(mx - M) for _, mx, sm in parts)
    merged = (num / den.unsqueeze(-1)).reshape(m, H * C)
    assert rel_err(merged, full) < FP32_TOL


def test_gat_host_buffer_entry_point():
    H, C, E, T = 4, 8, 500, 40
    target, XL, XR, att, bias, _ = _gat_case(E, T, H, C, seed=2)
    rng = np.random.default_rng(0)
    shuffle = rng.permutation(E)                       # host API takes edges in any order
    target, XL = target[shuffle], np.ascontiguousarray(XL[shuffle])
    out = np.zeros((T, H * C), np.float32)
    lib = _lib.load()
    tgt = np.ascontiguousarray(target, np.int64)
    rc = lib.gasfm_gat_edge_fwd_host(XL.ctypes.data, XR.ctypes.data, att.ctypes.data, bias.ctypes.data, tgt.ctypes.data,
                                     E, T, H, C, ctypes.c_float(0.2), out.ctypes.data)
    assert rc == 0, _lib.last_error()
    ref, _, _ = gat_edge_c.gat_edge_fwd(XL, XR, att, bias, target, T, H, C)
    assert rel_err(out, ref) < FP32_TOL


def test_pyg_style_forward_on_arbitrary_graph():
    from gasfm_b200.models.gatv2 import GATv2Conv
    from oracle.gatv2conv import GATv2Conv as RefConv
    torch.manual_seed(0)
    ours, ref = GATv2Conv(12, 8, heads=4, add_self_loops=False).to(DEV), RefConv(12, 8, heads=4, add_self_loops=False)
    ref.load_state_dict({k: v.cpu() for k, v in ours.state_dict().items()})
    x = torch.randn(50, 12)
    ei = torch.stack((torch.randint(0, 50, (300,)), torch.randint(0, 50, (300,))))
    want = ref(x.double() if False else x, ei)
    got = ours(x.to(DEV), ei.to(DEV))
    assert rel_err(got, want.detach().numpy()) < 5e-5


# ---------------------------------------------------------------------------------------------
# LayerNorm+ReLU, pooling, observation update
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w", [2, 32, 256, 12, 34, 1024])
@pytest.mark.parametrize("affine", [True, False])
def test_ln_relu_forward_backward(w, affine):
    torch.manual_seed(w)
    E = 777
    x = torch.randn(E, w, dtype=torch.float64) * 2 + 0.3
    gamma = (torch.randn(w, dtype=torch.float64) * 0.5 + 1).requires_grad_(True)
    beta = (torch.randn(w, dtype=torch.float64) * 0.2).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = torch.relu(torch.nn.functional.layer_norm(xr, (w,), gamma, beta, 1e-5)) if affine else torch.relu(xr)
    dy = torch.randn(E, w, dtype=torch.float64)
    ref.backward(dy)
    xg = x.float().to(DEV).requires_grad_(True)
    gg = gamma.detach().float().to(DEV).requires_grad_(True)
    bg = beta.detach().float().to(DEV).requires_grad_(True)
    y = ops.ln_relu(xg, gg, bg, 1e-5) if affine else ops.ln_relu(xg)
    y.backward(dy.float().to(DEV))
    assert rel_err(y, ref.detach().numpy()) < FP32_TOL
    # LayerNorm backward cancels catastrophically for tiny widths (w=2: xhat = +-1, dx ~ 0), so the
    # bound is "no worse than 3x the error of torch's own fp32 kernel against the same fp64 truth"
    xt = x.float().to(DEV).requires_grad_(True)
    yt = torch.relu(torch.nn.functional.layer_norm(xt, (w,), gamma.detach().float().to(DEV), beta.detach().float().to(DEV), 1e-5)) if affine else torch.relu(xt)
    yt.backward(dy.float().to(DEV))
    assert rel_err(xg.grad, xr.grad.numpy()) < max(GRAD_TOL, 3 * rel_err(xt.grad, xr.grad.numpy()))
    if affine:
        assert rel_err(gg.grad, gamma.grad.numpy()) < GRAD_TOL
        assert rel_err(bg.grad, beta.grad.numpy()) < GRAD_TOL


@pytest.mark.parametrize("w", [32, 256, 34])
@pytest.mark.parametrize("use", ["both", "skip_only", "main_only"])
def test_ln_relu_with_skip_sums_both_gradients_in_the_kernel(w, use):
    torch.manual_seed(w)
    E = 1501
    x = torch.randn(E, w, dtype=torch.float64)
    gamma = (torch.randn(w, dtype=torch.float64) * 0.5 + 1).requires_grad_(True)
    beta = (torch.randn(w, dtype=torch.float64) * 0.2).requires_grad_(True)
    c1, c2 = torch.randn(E, w, dtype=torch.float64), torch.randn(E, w, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    y_ref = torch.relu(torch.nn.functional.layer_norm(xr, (w,), gamma, beta, 1e-5))
    loss_ref = ((y_ref * c1).sum() if use != "skip_only" else 0) + ((xr * xr * c2).sum() if use != "main_only" else 0)
    loss_ref.backward()
    xg = x.float().to(DEV).requires_grad_(True)
    gg, bg = (t.detach().float().to(DEV).requires_grad_(True) for t in (gamma, beta))
    y, skip = ops.ln_relu_with_skip(xg, gg, bg, 1e-5)
    assert skip.data_ptr() == xg.data_ptr()
    c1g, c2g = c1.float().to(DEV), c2.float().to(DEV)
    loss = ((y * c1g).sum() if use != "skip_only" else 0) + ((skip * skip * c2g).sum() if use != "main_only" else 0)
    loss.backward()
    assert rel_err(y, y_ref.detach().numpy()) < FP32_TOL
    assert rel_err(xg.grad, xr.grad.numpy()) < GRAD_TOL
    if use != "skip_only":
        assert rel_err(gg.grad, gamma.grad.numpy()) < GRAD_TOL
        assert rel_err(bg.grad, beta.grad.numpy()) < GRAD_TOL


@pytest.mark.parametrize("rows,w,strided", [(1, 7, False), (1500, 256, False), (50000, 256, False), (20011, 33, True),
                                            (300, 64, False), (70000, 4, False)])
def test_col_sum_matches_fp64_sum(rows, w, strided):
    torch.manual_seed(rows + w)
    base = torch.randn(rows, w + (5 if strided else 0), dtype=torch.float32)
    x = base[:, :w]
    want = x.double().sum(dim=0).numpy()
    old = ops.COL_SUM_MIN_ROWS
    ops.COL_SUM_MIN_ROWS = 0          # force the kernel for the small cases too
    try:
        got = ops.col_sum(base.to(DEV)[:, :w])
        got_keep = ops.col_sum(base.to(DEV)[:, :w], keepdim=True)
    finally:
        ops.COL_SUM_MIN_ROWS = old
    scale = np.abs(x.double().numpy()).sum(axis=0).max()
    assert np.abs(got.cpu().numpy() - want).max() < 1e-6 * scale
    assert got_keep.shape == (1, w) and torch.equal(got_keep[0], got)


@pytest.mark.parametrize("K,N,bias", [(2, 256, True), (1, 32, True), (4, 64, False), (3, 1024, True), (2, 36, True)])
def test_linear_with_tiny_input_width_uses_the_write_bound_kernel(K, N, bias):
    """first block: the observations are 2 wide, so lin_l / lin_proj / skip_projection are [E,2] x [2,N]"""
    torch.manual_seed(K * N)
    E = 5003
    x = torch.randn(E, K, dtype=torch.float64, requires_grad=True)
    W = torch.randn(N, K, dtype=torch.float64, requires_grad=True)
    b = torch.randn(N, dtype=torch.float64, requires_grad=True) if bias else None
    dy = torch.randn(E, N, dtype=torch.float64)
    torch.nn.functional.linear(x, W, b).backward(dy)
    xg, Wg = (t.detach().float().to(DEV).requires_grad_(True) for t in (x, W))
    bg = b.detach().float().to(DEV).requires_grad_(True) if bias else None
    before = _lib.launch_count
    y = ops.linear(xg, Wg, bg)
    if K < 4:      # K = 4 is also a valid tensor-core shape (weight split + GEMM)
        assert _lib.launch_count == before + 1, "expected exactly one gasfm kernel call (no cuBLAS fallback)"
    y.backward(dy.float().to(DEV))
    assert rel_err(y, torch.nn.functional.linear(x, W, b).detach().numpy()) < FP32_TOL
    assert rel_err(xg.grad, x.grad.numpy()) < GRAD_TOL
    assert rel_err(Wg.grad, W.grad.numpy()) < GRAD_TOL
    if bias:
        assert rel_err(bg.grad, b.grad.numpy()) < GRAD_TOL


@pytest.mark.parametrize("w", [6, 32, 256, 3, 100])
def test_row_col_pooling_matches_reference_golden_and_oracle(w):
    g = load_golden("pooling")
    idx = torch.from_numpy(g["indices"])
    m, n = int(g["shape"][0]), int(g["shape"][1])
    if w == 6:
        feat = torch.from_numpy(g["feat"])
    else:
        feat = torch.randn(idx.shape[1], w)
    cam_per_pts = torch.bincount(idx[1], minlength=n).unsqueeze(1).to(DEV)
    pts_per_cam = torch.bincount(idx[0], minlength=m).unsqueeze(1).to(DEV)
    sm = sparse_utils.SparseMat(feat.to(DEV).requires_grad_(True), idx.to(DEV), cam_per_pts, pts_per_cam, (m, n, w))
    s0, s1 = sm.sum(0), sm.sum(1)
    assert rel_err(s0, gasfm_cpu.sparse_sum(feat.double(), idx, (m, n, w), 0).numpy()) < FP32_TOL
    assert rel_err(s1, gasfm_cpu.sparse_sum(feat.double(), idx, (m, n, w), 1).numpy()) < FP32_TOL
    if w == 6:
        np.testing.assert_allclose(s0.detach().cpu().numpy(), g["sum0"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(s1.detach().cpu().numpy(), g["sum1"], rtol=1e-5, atol=1e-5)
        m0 = sm.mean(0).detach().cpu().numpy()
        assert np.array_equal(np.isnan(m0), np.isnan(g["mean0"]))      # 0/0 for empty tracks, as in the reference
        np.testing.assert_allclose(np.nan_to_num(m0), np.nan_to_num(g["mean0"]), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(sm.mean(1).detach().cpu().numpy(), g["mean1"], rtol=1e-5, atol=1e-5)
    # backward of pooling = broadcast
    (s0 * 2.0).sum().backward()
    assert rel_err(sm.values.grad, np.full((idx.shape[1], w), 2.0)) < 1e-6
    mean_cols = sparse_utils.sparse_mean(sm, 0)
    cnt = np.maximum(np.bincount(g["indices"][1], minlength=n), 1)[:, None]
    assert rel_err(mean_cols, gasfm_cpu.sparse_sum(feat.double(), idx, (m, n, w), 0).numpy() / cnt) < FP32_TOL


def test_set_of_set_layer_matches_reference_golden():
    from gasfm_b200.models.layers import SetOfSetLayer
    g = load_golden("pooling")
    layer = SetOfSetLayer(6, 10)
    layer.load_state_dict({k[len("param."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")})
    idx = torch.from_numpy(g["indices"])
    sm = sparse_utils.SparseMat(torch.from_numpy(g["feat"]).to(DEV), idx.to(DEV),
                                torch.bincount(idx[1], minlength=50).unsqueeze(1).to(DEV),
                                torch.bincount(idx[0], minlength=7).unsqueeze(1).to(DEV), (7, 50, 6))
    out = layer.to(DEV)(sm)
    np.testing.assert_allclose(out.values.detach().cpu().numpy(), g["sos_out"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("w,d0", [(32, 2), (256, 2), (20, 0), (16, 3), (6, 2)])
def test_edge_update_forward_backward(w, d0):
    m, n = 11, 150
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, 700, seed=w)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m, n)
    torch.manual_seed(1)
    mk = lambda *s: torch.randn(*s, dtype=torch.float64).requires_grad_(True)  # noqa: E731
    P, S, V, gl, skip = mk(E, w), mk(n, w), mk(m, w), mk(1, w), mk(E, w)
    x0, W0 = (mk(E, d0), mk(w, d0)) if d0 else (None, None)
    r, c = torch.from_numpy(idx_np[0]), torch.from_numpy(idx_np[1])
    ref = P + 0.25 * (S[c] + V[r] + gl + (x0 @ W0.T if d0 else 0)) + skip
    dout = torch.randn(E, w, dtype=torch.float64)
    ref.backward(dout)
    f = lambda t: None if t is None else t.detach().float().to(DEV).requires_grad_(True)  # noqa: E731
    Pg, Sg, Vg, gg, sg, x0g, W0g = map(f, (P, S, V, gl, skip, x0, W0))
    out = ops.edge_update(Pg, x0g, W0g, Sg, Vg, gg, sg, oi, 1.0, 0.25)
    out.backward(dout.float().to(DEV))
    assert rel_err(out, ref.detach().numpy()) < FP32_TOL
    for got, want in ((Pg, P), (Sg, S), (Vg, V), (gg, gl), (sg, skip), (x0g, x0), (W0g, W0)):
        if want is not None:
            assert rel_err(got.grad, want.grad.numpy()) < GRAD_TOL


# ---------------------------------------------------------------------------------------------
# tcgen05 3xTF32 GEMM
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (1000, 256, 256), (4097, 32, 32), (300, 64, 100), (20000, 256, 260),
                                   (5000, 48, 36), (129, 256, 4), (70000, 128, 64)])
def test_gemm_tf32x3_has_fp32_accuracy(M, N, K):
    """C = A W^T + b on the tensor cores must keep fp32-level accuracy (the reference's GEMMs are true
    fp32): within 5x of cuBLAS fp32's own error against fp64, and ~100x better than single-pass TF32."""
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    b = torch.randn(N, device=DEV)
    ref = a.double() @ w.double().t() + b.double()
    got = ops.gemm_tf32x3(a, w, b)
    err = ((got.double() - ref).abs().max() / ref.abs().max()).item()
    err32 = ((torch.nn.functional.linear(a, w, b).double() - ref).abs().max() / ref.abs().max()).item()
    assert err < max(5 * err32, 4e-6), (err, err32)
    # strided A (a column slice of a wider matrix) and no bias
    wide = torch.randn(M, K + 8, device=DEV)
    got2 = ops.gemm_tf32x3(wide[:, 4:4 + K], w)
    ref2 = wide[:, 4:4 + K].double() @ w.double().t()
    assert ((got2.double() - ref2).abs().max() / ref2.abs().max()).item() < max(5 * err32, 4e-6)


@pytest.mark.parametrize("M,N,K", [(128, 16, 64), (1000, 256, 256), (4097, 32, 72), (300, 64, 104), (20000, 256, 256),
                                   (5000, 48, 80), (129, 256, 128), (70000, 128, 64), (70000, 64, 256), (33000, 256, 192)])
def test_gemm_f16x2_has_fp32_accuracy_for_any_row_scale(M, N, K):
    """The scaled 2 x FP16 split must keep fp32-level accuracy whatever the magnitude of a row: rows of A spanning
    1e-12 .. 1e+12 (gradients are tiny, pre-norm features can be large), weights rows likewise, zero rows, and a
    row whose entries span 7 decades.  Error is measured per ROW against that row's own largest output."""
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV)
    a *= 10.0 ** torch.randint(-12, 13, (M, 1), device=DEV).float()        # per-row magnitude
    a[::97] = 0.0
    a[1] = torch.randn(K, device=DEV) * 10.0 ** torch.linspace(-7, 0, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    w *= 10.0 ** torch.randint(-3, 4, (N, 1), device=DEV).float()
    b = None
    ref = a.double() @ w.double().t()
    assert ops.gemm_f16x2_supported(M, N, K, K, N)
    got = ops.gemm_f16x2(a, w, b)
    # per-row, per-column normalisation: |err[m,n]| <= tol * |a_m| |w_n|
    bound = a.double().norm(dim=1, keepdim=True) * w.double().norm(dim=1).unsqueeze(0) + 1e-300
    err = ((got.double() - ref).abs() / bound).max().item()
    err32 = ((torch.nn.functional.linear(a, w).double() - ref).abs() / bound).max().item()
    assert err < max(5 * err32, 2e-6), (err, err32)
    assert torch.isfinite(got).all()
    # bias, strided A, accumulate mode
    bias = torch.randn(N, device=DEV)
    wide = torch.randn(M, K + 8, device=DEV)
    base = torch.randn(M, N, device=DEV)
    out = base.clone()
    got2 = ops.gemm_f16x2(wide[:, 4:4 + K], w, bias, out=out, accumulate=True)
    assert got2.data_ptr() == out.data_ptr()
    ref2 = base.double() + wide[:, 4:4 + K].double() @ w.double().t() + bias.double()
    assert ((got2.double() - ref2).abs().max() / ref2.abs().max()).item() < 4e-6


def test_gemm_f16x2_groups_read_the_input_once():
    """three projections of one input in one launch == three separate GEMMs, bit for bit"""
    torch.manual_seed(11)
    M, K, N = 30011, 256, 256
    a = torch.randn(M, K, device=DEV) * 10.0 ** torch.randint(-4, 5, (M, 1), device=DEV).float()
    ws = [torch.randn(N, K, device=DEV) / K ** 0.5 for _ in range(3)]
    bs = [torch.randn(N, device=DEV), None, torch.randn(N, device=DEV)]
    out = ops.gemm_f16x2_groups(a, ws, bs)
    assert out.shape == (M, 3 * N)
    for g in range(3):
        assert torch.equal(out[:, g * N:(g + 1) * N], ops.gemm_f16x2(a, ws[g], bs[g]))
    assert not ops.gemm_f16x2_supported(M, N, 32, 32, N)      # narrow K stays on the 3xTF32 kernel


def test_linear_autograd_matches_fp64():
    torch.manual_seed(3)
    M, N, K = 6000, 64, 32
    x = torch.randn(M, K, device=DEV, requires_grad=True)
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, device=DEV, requires_grad=True)
    dy = torch.randn(M, N, device=DEV)
    y = ops.linear(x, w, b)
    assert y.grad_fn is not None and type(y.grad_fn).__name__.startswith("_LinearTC")
    y.backward(dy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    (torch.nn.functional.linear(xd, wd, bd) * dy.double()).sum().backward()
    assert rel_err(y, torch.nn.functional.linear(xd, wd, bd).detach().cpu().numpy()) < FP32_TOL
    assert rel_err(x.grad, xd.grad.cpu().numpy()) < FP32_TOL
    assert rel_err(w.grad, wd.grad.cpu().numpy()) < GRAD_TOL
    assert rel_err(b.grad, bd.grad.cpu().numpy()) < GRAD_TOL


@pytest.mark.parametrize("E,Nout,Kout", [(16, 32, 32), (1000, 256, 256), (4099, 32, 32), (3000, 64, 48), (70001, 256, 32),
                                         (300000, 32, 256), (17, 4, 16)])
def test_wgrad_tf32x3_has_fp32_accuracy(E, Nout, Kout):
    """dW = dY^T X on the tensor cores (MN-major operands, split-K over SMs, accumulator drained every
    256 stages): error against fp64 within 5x of cuBLAS fp32's, never worse than 1e-5 of the largest entry."""
    torch.manual_seed(E + Nout)
    dy = torch.randn(E, Nout, device=DEV)
    x = torch.randn(E, Kout, device=DEV)
    ref = dy.double().t() @ x.double()
    got = ops.wgrad_tf32x3(dy, x)
    err = ((got.double() - ref).abs().max() / ref.abs().max()).item()
    err32 = (((dy.t() @ x).double() - ref).abs().max() / ref.abs().max()).item()
    assert err < max(5 * err32, 1e-5), (err, err32)
    assert torch.equal(got, ops.wgrad_tf32x3(dy, x))           # deterministic split-K


@pytest.mark.parametrize("E,Nout,Kout", [(5, 32, 32), (4099, 32, 32), (30000, 64, 32), (30001, 32, 64), (9000, 64, 64)])
def test_wgrad_small_widths(E, Nout, Kout):
    """shipped widths (32 / 64): SIMT register-tiled dW, plain fp32 accuracy, deterministic"""
    torch.manual_seed(E)
    wide = torch.randn(E, Nout + 8, device=DEV)
    dy, x = wide[:, 4:4 + Nout], torch.randn(E, Kout, device=DEV)      # strided dY rows
    assert _lib.load().gasfm_wgrad_small_supported(Nout, Kout, Nout + 8, Kout)
    got = ops.wgrad_tf32x3(dy, x)
    ref = dy.double().t() @ x.double()
    assert rel_err(got, ref.cpu().numpy()) < 2e-6
    assert torch.equal(got, ops.wgrad_tf32x3(dy, x))


def test_full_size_edge_kernels_cfg3():
    """BASELINE.json configs[2] size (1000 views x 300k tracks, ~5M observations) at the shipped width:
    size-independent properties of the edge kernels at full size."""
    m, n, H, C = 1000, 300000, 4, 8
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, 5_000_000, seed=0)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m, n)
    torch.manual_seed(0)
    XL = torch.randn(E, H * C, device=DEV)
    att = torch.randn(1, H, C, device=DEV) * 0.3
    for plan, T in ((oi.by_view, m), (oi.by_track, n)):
        XR = torch.randn(T, H * C, device=DEV)
        const = torch.arange(H * C, device=DEV, dtype=torch.float32).repeat(E, 1) * 0.1
        out = ops.gat_edge_attention(const, XR, att, None, plan, H)
        assert torch.allclose(out, const[:1].expand(T, -1), atol=2e-5, rtol=1e-5)     # weights sum to one
        xl, xr = XL.clone().requires_grad_(True), XR.clone().requires_grad_(True)
        o = ops.gat_edge_attention(xl, xr, att, None, plan, H)
        o.sum().backward()
        # with dOut = 1:  sum over a segment of dXL = sum_e alpha_e + sum_e dz_e = 1 + dXR[t]
        seg_tot = ops.seg_sum_raw(xl.grad, plan)
        nonempty = (plan.seg_ptr[1:] > plan.seg_ptr[:-1]).unsqueeze(1)
        assert torch.allclose(seg_tot, torch.where(nonempty, 1.0 + xr.grad, torch.zeros_like(seg_tot)), atol=5e-4)
    assert torch.equal(oi.csc_perm.long(), torch.argsort(torch.from_numpy(idx_np[1]).to(DEV), stable=True))


@pytest.mark.parametrize("M,K,widths", [(9000, 64, (64, 32, 48)), (9000, 64, (32, 32, 32)), (70000, 256, (256, 256, 256)),
                                        (5000, 32, (64, 64))])
def test_linear_multi_sums_input_gradients_in_one_pass(M, K, widths):
    """several projections of one input (lin_l x2 + lin_proj): outputs, dX (equal widths: ONE GEMM over the
    concatenated [dY_0 | dY_1 | ..]; otherwise accumulated in the GEMM epilogues), dW and the fused bias gradients
    against fp64"""
    torch.manual_seed(5)
    x = torch.randn(M, K, device=DEV, requires_grad=True)
    ws = [(torch.randn(n, K, device=DEV) / K ** 0.5).requires_grad_(True) for n in widths]
    bs = [torch.randn(n, device=DEV, requires_grad=True) for n in widths]
    ys = ops.linear_multi(x, list(zip(ws, bs)))
    assert type(ys[0].grad_fn).__name__.startswith("_LinearMulti")
    dys = [torch.randn_like(y) for y in ys]
    torch.autograd.backward(ys, dys)
    xd = x.detach().double().requires_grad_(True)
    wd = [w.detach().double().requires_grad_(True) for w in ws]
    bd = [b.detach().double().requires_grad_(True) for b in bs]
    yd = [torch.nn.functional.linear(xd, w, b) for w, b in zip(wd, bd)]
    torch.autograd.backward(yd, [d.double() for d in dys])
    for y, r in zip(ys, yd):
        assert rel_err(y, r.detach().cpu().numpy()) < FP32_TOL
    assert rel_err(x.grad, xd.grad.cpu().numpy()) < FP32_TOL
    for w, r in zip(ws, wd):
        assert rel_err(w.grad, r.grad.cpu().numpy()) < GRAD_TOL
    for b, r in zip(bs, bd):
        assert rel_err(b.grad, r.grad.cpu().numpy()) < GRAD_TOL


@pytest.mark.parametrize("E,Nout,Kout", [(64, 128, 64), (1000, 256, 256), (4099, 128, 192), (70001, 256, 64), (200003, 256, 256)])
@pytest.mark.parametrize("slack", [1.0, 37.0])
def test_wgrad_f16x2_has_fp32_accuracy(E, Nout, Kout, slack):
    """dW = dY^T X on the fp16 tensor-core path (MN-major operands written by the producer warps, one power-of-two scale
    per operand from the matrix maximum): error against fp64 no worse than the 3xTF32 kernel's bound (1e-5 of the
    largest entry), with rows of dY spanning 5 decades and the maxima overestimated by ``slack`` (any bound works)."""
    torch.manual_seed(E + Nout)
    dy = torch.randn(E, Nout, device=DEV) * 10.0 ** torch.randint(-5, 1, (E, 1), device=DEV).float() * 1e-3
    x = torch.relu(torch.randn(E, Kout, device=DEV) * 3)
    assert ops.wgrad_f16x2_supported(E, Nout, Kout, Nout, Kout)
    ady, ax = (dy.abs().max() * slack).reshape(1), (x.abs().max() * slack).reshape(1)
    dw, db = ops.wgrad_f16x2(dy, x, ady, ax, with_bias=True)
    ref = dy.double().t() @ x.double()
    assert ((dw.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
    refb = dy.double().sum(0)
    assert ((db.double() - refb).abs().max() / refb.abs().max()).item() < 2e-6
    # the maxima the GEMMs leave behind are the true ones
    w = torch.randn(64, Kout, device=DEV)
    _, amax_x = ops.gemm_f16x2_groups(x, [w, w], [None, None], want_amax=True) if Kout >= 64 else (None, ax)
    assert torch.equal(amax_x, x.abs().max().reshape(1))
    _, amax_cat = ops.gemm_tf32x3_cat([dy, 2 * dy], torch.randn(32, 2 * Nout, device=DEV), want_amax=True)
    assert torch.equal(amax_cat, torch.stack([dy.abs().max(), (2 * dy).abs().max()]))


@pytest.mark.parametrize("E,Nout,Kout", [(70001, 256, 256), (5000, 48, 80), (30000, 64, 64), (999, 32, 32)])
def test_wgrad_fused_bias_gradient(E, Nout, Kout):
    torch.manual_seed(E)
    dy, x = torch.randn(E, Nout, device=DEV) + 0.3, torch.randn(E, Kout, device=DEV)
    dw, db = ops.wgrad_tf32x3(dy, x, with_bias=True)
    assert rel_err(db, dy.double().sum(0).cpu().numpy()) < 2e-6
    assert torch.equal(dw, ops.wgrad_tf32x3(dy, x))


@pytest.mark.parametrize("E,n_groups", [(70001, 3), (5000, 2), (200003, 3), (40, 3)])
def test_wgrad_f16x2_multi_equals_separate_launches(E, n_groups):
    """The weight gradients of a block's projections in ONE launch (groups share the reads of x through L2): bit-identical
    to one launch per projection when the row split is the same, fp32-accurate against fp64 otherwise."""
    torch.manual_seed(E)
    x = torch.relu(torch.randn(E, 256, device=DEV))
    dys = [torch.randn(E + 3, 256, device=DEV)[3:] * (10.0 ** -g) for g in range(n_groups)]       # distinct buffers and scales
    amax_x = x.abs().max().reshape(1)
    amax = torch.stack([dy.abs().max() for dy in dys])
    dw, db = ops.wgrad_f16x2_multi(dys, x, amax, amax_x)
    assert dw.shape == (n_groups, 256, 256) and db.shape == (n_groups, 256)
    for g, dy in enumerate(dys):
        ref = dy.double().t() @ x.double()
        assert ((dw[g].double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
        refb = dy.double().sum(0)
        assert ((db[g].double() - refb).abs().max() / refb.abs().max()).item() < 2e-6
    dw2, db2 = ops.wgrad_f16x2_multi(dys, x, amax, amax_x)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)                 # deterministic


@pytest.mark.parametrize("M,K,N,G", [(1000, 256, 256, 3), (70001, 256, 256, 2), (4097, 128, 64, 3), (300, 64, 32, 1), (129, 192, 256, 2)])
def test_gemm_f16x2_with_fused_layernorm_relu(M, K, N, G):
    """LayerNorm + ReLU evaluated inside the GEMM's operand producer (x_raw read once, relu(LN(x)) never in memory):
    projections, row statistics and the operand maximum against fp64; rows with very different scales."""
    torch.manual_seed(M + K)
    x = torch.randn(M, K, device=DEV) * (10.0 ** torch.randint(-3, 3, (M, 1), device=DEV).float()) + 0.3
    gamma, beta = torch.rand(K, device=DEV) + 0.5, torch.randn(K, device=DEV) * 0.3
    ws = [torch.randn(N, K, device=DEV) / K ** 0.5 for _ in range(G)]
    bs = [torch.randn(N, device=DEV) for _ in range(G)]
    out, amax, mean, rstd = ops.gemm_f16x2_groups_ln(x, gamma, beta, 1e-5, ws, bs)
    xd = x.double()
    mu, var = xd.mean(1, keepdim=True), xd.var(1, unbiased=False, keepdim=True)
    y = torch.relu((xd - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double())
    assert rel_err(mean, mu[:, 0].cpu().numpy()) < 1e-5 * max(1.0, float(mu.abs().max()))
    assert float(((rstd.double() - 1 / torch.sqrt(var[:, 0] + 1e-5)).abs() * torch.sqrt(var[:, 0] + 1e-5)).max()) < 1e-5
    assert abs(float(amax) - float(y.max())) < 1e-5 * float(y.max())
    for g in range(G):
        ref = y @ ws[g].double().t() + bs[g].double()
        assert rel_err(out[:, g * N:(g + 1) * N], ref.cpu().numpy()) < FP32_TOL
    # the unfused pair of kernels gives the same projections to fp32 accuracy
    y32 = ops.ln_relu(x, gamma, beta, 1e-5)
    ref32 = ops.gemm_f16x2_groups(y32, ws, bs) if G > 1 else ops.gemm_f16x2(y32, ws[0], bs[0])
    assert rel_err(out, ref32.double().cpu().numpy()) < FP32_TOL
    # N = K = 256 (CTA-pair kernel): the normalised operand as a by-product of the same pass, projections bit-identical
    assert ops.gemm_f16x2_ln_y_supported(M, N, K, K, G * N) == (N == 256 and K == 256)
    if N == 256 and K == 256:
        out2, amax2, mean2, rstd2, y_out = ops.gemm_f16x2_groups_ln(x, gamma, beta, 1e-5, ws, bs, want_y=True)
        assert torch.equal(out2, out) and torch.equal(mean2, mean) and torch.equal(rstd2, rstd) and torch.equal(amax2, amax)
        assert rel_err(y_out, y.cpu().numpy()) < 1e-5 and float(y_out.min()) >= 0.0


@pytest.mark.parametrize("M,N,n_seg,seg_k", [(70001, 256, 3, 256), (1000, 256, 2, 256), (4099, 64, 3, 128), (129, 256, 4, 256), (300, 32, 1, 128)])
def test_gemm_f16x2_cat_has_fp32_accuracy(M, N, n_seg, seg_k):
    """The concatenated input gradient on the fp16 path: [A_0 | A_1 | ..] W^T with ONE row scale across the segments (two-pass
    operand producer), segments and rows of very different magnitude, against fp64; per-segment maxima exact."""
    torch.manual_seed(M + n_seg)
    row_scale = 10.0 ** torch.randint(-4, 4, (M, 1), device=DEV).float()
    segs = [torch.randn(M + 2, seg_k, device=DEV)[2:] * row_scale * (10.0 ** (-2 * i)) for i in range(n_seg)]
    w = torch.randn(N, n_seg * seg_k, device=DEV) / (n_seg * seg_k) ** 0.5
    bias = torch.randn(N, device=DEV)
    got, amax = ops.gemm_f16x2_cat(segs, w, None, want_amax=True)
    ref = torch.cat([a.double() for a in segs], dim=1) @ w.double().t()
    # error relative to the size of the row's products (|a_row| |w_row|), like the forward kernel's test (no bias here: for the
    # tiny rows the fp32 rounding of acc + bias would dominate any kernel's error)
    scale = torch.cat(segs, dim=1).double().norm(dim=1, keepdim=True) * w.double().norm(dim=1).unsqueeze(0) + 1e-30
    assert float(((got.double() - ref).abs() / scale).max()) < 2e-6
    assert torch.equal(amax, torch.stack([a.abs().max() for a in segs]))
    ref32 = ops.gemm_tf32x3_cat(segs, w, None)
    assert float(((got - ref32).double().abs() / scale).max()) < 4e-6
    assert torch.equal(got, ops.gemm_f16x2_cat(segs, w, None))
    with_bias = ops.gemm_f16x2_cat(segs, w, bias)
    assert float((with_bias.double() - (ref + bias.double())).abs().max() / (ref + bias.double()).abs().max()) < 1e-5


def test_row_maxima_from_upstream_kernels_feed_the_one_pass_fp16_input_gradient():
    """The attention backward and the update's x0 kernel emit max |row| of the gradients they produce; with those the fp16
    concatenated GEMM runs its one-pass producer and must give bit-identical results to its two-pass form."""
    H, C, m, n = 4, 64, 12, 700
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, 6000, seed=9)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m, n)
    torch.manual_seed(1)
    XL = torch.randn(E, H * C, device=DEV)
    att = torch.randn(1, H, C, device=DEV) * 0.3
    grads, rowmaxes = [], []
    for plan, T in ((oi.by_track, n), (oi.by_view, m)):
        XR = torch.randn(T, H * C, device=DEV)
        acc, mx, sm = ops.gat_edge_partial(XL, XR, att, plan, H)
        out = acc / sm.repeat_interleave(C, dim=1).clamp_min(1e-30)
        dO = torch.randn(T, H * C, device=DEV) * 10.0 ** torch.randint(-3, 3, (T, 1), device=DEV).float()
        dXL, _, _, rm = ops.gat_edge_backward_raw(XL, XR, att, out, mx, sm, dO, plan, H, want_rowmax=True)
        ref, _, _ = ops.gat_edge_backward_raw(XL, XR, att, out, mx, sm, dO, plan, H)
        assert torch.equal(dXL, ref) and torch.equal(rm, dXL.abs().amax(dim=1))
        grads.append(dXL); rowmaxes.append(rm)
    d_out = torch.randn(E, H * C, device=DEV) * 10.0 ** torch.randint(-4, 2, (E, 1), device=DEV).float()
    x0, W0 = torch.randn(E, 2, device=DEV), torch.randn(H * C, 2, device=DEV)
    dx0, dW0, rm = ops._x0_backward(d_out, x0, W0, 0.25, want_rowmax=True)
    dx0_ref, dW0_ref, _ = ops._x0_backward(d_out, x0, W0, 0.25)
    assert torch.equal(dx0, dx0_ref) and torch.equal(dW0, dW0_ref) and torch.equal(rm, d_out.abs().amax(dim=1))
    grads.append(d_out); rowmaxes.append(rm)
    w = torch.randn(H * C, 3 * H * C, device=DEV) / (3 * H * C) ** 0.5
    one_pass, amax1 = ops.gemm_f16x2_cat(grads, w, want_amax=True, rowmax=rowmaxes)
    two_pass, amax2 = ops.gemm_f16x2_cat(grads, w, want_amax=True)
    assert torch.equal(one_pass, two_pass) and torch.equal(amax1, amax2)
    ref = torch.cat([g.double() for g in grads], dim=1) @ w.double().t()
    scale = torch.cat(grads, dim=1).double().norm(dim=1, keepdim=True) * w.double().norm(dim=1).unsqueeze(0) + 1e-30
    assert float(((one_pass.double() - ref).abs() / scale).max()) < 2e-6


@pytest.mark.parametrize("w,d0", [(256, 2), (32, 2), (64, 3), (128, 1)])
def test_update_backward_views_fused_pass(w, d0):
    """One storage-order pass over dOut for the update's backward: per-view sums, the x0 term's gradients and the row maxima
    must equal the separate kernels (segment sums up to summation order), views without observations included."""
    m, n = 14, 900
    idx_np, _ = gasfm_cpu.synthetic_observations(m, n, 9000, seed=w + d0)
    E = idx_np.shape[1]
    oi = ObservationIndex(torch.from_numpy(idx_np).to(DEV), m + 2, n)      # two trailing views without observations
    assert oi.by_view.chunk > 0
    torch.manual_seed(w)
    d_out = torch.randn(E, w, device=DEV) * 10.0 ** torch.randint(-3, 2, (E, 1), device=DEV).float()
    x0, W0 = torch.randn(E, d0, device=DEV), torch.randn(w, d0, device=DEV)
    dV, dx0, dW0, rm = ops._update_backward_views(d_out, x0, W0, 0.25, oi.by_view, want_rowmax=True)
    dV_ref = ops.seg_sum_raw(d_out, oi.by_view, 0.25)
    dx0_ref, dW0_ref, rm_ref = ops._x0_backward(d_out, x0, W0, 0.25, want_rowmax=True)
    assert torch.equal(rm, rm_ref) and torch.equal(dx0, dx0_ref)
    assert rel_err(dV, dV_ref.double().cpu().numpy()) < 2e-6 and torch.equal(dV[-2:], torch.zeros(2, w, device=DEV))
    ref_w = 0.25 * d_out.double().t() @ x0.double()
    assert rel_err(dW0, ref_w.cpu().numpy()) < 5e-6 and rel_err(dW0_ref, ref_w.cpu().numpy()) < 5e-6
    assert rel_err(dV, (0.25 * torch.zeros(m + 2, w, dtype=torch.float64, device=DEV).index_add_(0, oi.row_idx.long(), d_out.double())).cpu().numpy()) < 2e-6
