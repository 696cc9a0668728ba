"""The CPU oracle against the golden fixtures that were generated from the unmodified reference
(tests/golden/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import numpy as np
import pytest
import torch

from conftest import MODEL_FIXTURES, golden_model, grad_errors, load_golden
from oracle import gasfm_cpu, gat_edge_c
from oracle.gatv2conv import gatv2_edge_softmax_aggregate


def test_index_build_bit_exact():
    g = load_golden("index_build")
    scene = gasfm_cpu.make_scene(torch.from_numpy(g["M"]), torch.from_numpy(g["Ns"]))
    x = scene["x"]
    assert np.array_equal(x["indices"].numpy(), g["indices"])
    assert np.array_equal(x["cam_per_pts"].numpy(), g["cam_per_pts"])
    assert np.array_equal(x["pts_per_cam"].numpy(), g["pts_per_cam"])
    assert tuple(x["shape"]) == tuple(g["shape"])
    np.testing.assert_allclose(x["values"].numpy(), g["values"], rtol=1e-6, atol=1e-7)
    for name, gr in scene["graphs"].items():
        assert np.array_equal(gr["edge_index"].numpy(), g[f"{name}.edge_index"]), name
        assert np.array_equal(gr["valid_indices"].numpy(), g[f"{name}.valid_indices"]), name
        assert [gr["m"], gr["n"], gr["agg_dim"], gr["n_agg_nodes"]] == list(g[f"{name}.meta"]), name
    # degenerate tracks of the fixture: never observed / observed once -> dropped
    assert g["cam_per_pts"][3] == 0 and g["cam_per_pts"][5] == 0
    # the view with < 8 points is not a source of view2global
    assert 6 not in g["view2global.valid_indices"][0]


def test_pooling_and_set_of_set_layer():
    g = load_golden("pooling")
    feat, idx = torch.from_numpy(g["feat"]), torch.from_numpy(g["indices"])
    shape = tuple(g["shape"])
    np.testing.assert_allclose(gasfm_cpu.sparse_sum(feat, idx, shape, 0).numpy(), g["sum0"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gasfm_cpu.sparse_sum(feat, idx, shape, 1).numpy(), g["sum1"], rtol=1e-6, atol=1e-6)
    m0 = gasfm_cpu.sparse_mean(feat, idx, shape, 0).numpy()
    assert np.array_equal(np.isnan(m0), np.isnan(g["mean0"]))          # empty tracks: 0/0 like the reference
    np.testing.assert_allclose(np.nan_to_num(m0), np.nan_to_num(g["mean0"]), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gasfm_cpu.sparse_mean(feat, idx, shape, 1).numpy(), g["mean1"], rtol=1e-6, atol=1e-6)
    params = {k[len("param."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")}
    out = gasfm_cpu.set_of_set_layer(gasfm_cpu._P(params), feat, idx, shape)
    np.testing.assert_allclose(out.numpy(), g["sos_out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", MODEL_FIXTURES)
@pytest.mark.parametrize("tag,dtype,tol,gtol", [("f32", torch.float32, 2e-5, 2e-3), ("f64", torch.float64, 1e-11, 1e-8)])
def test_model_matches_reference_outputs_and_grads(name, tag, dtype, tol, gtol):
    conf, params, g = golden_model(name)
    mc = conf["model"]
    params = {k: v.to(dtype).requires_grad_(True) for k, v in params.items()}
    scene = gasfm_cpu.make_scene(torch.from_numpy(g["M"]).to(dtype), torch.from_numpy(g["Ns"]).to(dtype))
    out = gasfm_cpu.gasfm_forward(params, scene, n_heads=mc["n_heads"], stateful=mc["stateful_global_features"],
                                  calibrated=conf["dataset"]["calibrated"],
                                  rot_representation=mc["view_head"].get("rot_representation", "quat"),
                                  normalize_output=mc["view_head"].get("normalize_output"))
    for key in ("Ps_norm", "pts3D"):
        want = g[f"out.{tag}.{key}"]
        err = np.abs(out[key].detach().numpy() - want).max()
        assert err <= tol * max(1.0, np.abs(want).max()), (key, err)
    w = torch.linspace(0.5, 1.5, out["Ps_norm"].numel(), dtype=dtype).reshape(out["Ps_norm"].shape)
    w2 = torch.linspace(-1.0, 1.0, out["pts3D"].numel(), dtype=dtype).reshape(out["pts3D"].shape)
    loss = (out["Ps_norm"] * w).sum() + (out["pts3D"] * w2).sum()
    if "depths" in out:
        np.testing.assert_allclose(out["depths"].detach().numpy(), g[f"out.{tag}.depths"], rtol=0, atol=tol * 10)
        loss = loss + (out["depths"] ** 2).sum()
    loss.backward()
    want = {k[len(f"grad.{tag}."):]: g[k] for k in g.files if k.startswith(f"grad.{tag}.")}
    got = {k: params[k].grad.numpy() for k in want}
    worst, key = grad_errors(got, want)
    assert worst < gtol, (key, worst)


@pytest.mark.parametrize("H,C", [(4, 8), (4, 1), (2, 6), (4, 64)])
def test_c_oracle_agrees_with_torch_restatement(H, C):
    """Two independent restatements of the GATv2 edge arithmetic (plain C fp64, torch) agree,
    including empty segments, and so does the hand-derived backward with torch autograd."""
    rng = np.random.default_rng(H * 100 + C)
    E, T = 200, 23
    target = rng.integers(0, T - 3, size=E)          # last 3 targets have no edges
    XL = rng.standard_normal((E, H * C)).astype(np.float32)
    XR = rng.standard_normal((T, H * C)).astype(np.float32)
    att = rng.standard_normal(H * C).astype(np.float32)
    bias = rng.standard_normal(H * C).astype(np.float32)
    out, smax, ssum = gat_edge_c.gat_edge_fwd(XL, XR, att, bias, target, T, H, C)
    xl = torch.cat((torch.from_numpy(XL).double(), torch.zeros(T, H * C, dtype=torch.float64))).requires_grad_(True)
    xr = torch.cat((torch.zeros(E, H * C, dtype=torch.float64), torch.from_numpy(XR).double())).requires_grad_(True)
    a = torch.from_numpy(att).double().reshape(1, H, C).requires_grad_(True)
    ei = torch.stack((torch.arange(E), E + torch.from_numpy(target)))
    o = gatv2_edge_softmax_aggregate(xl.view(-1, H, C), xr.view(-1, H, C), a, ei).reshape(-1, H * C)[E:] + torch.from_numpy(bias).double()
    np.testing.assert_allclose(out, o.detach().numpy(), rtol=1e-12, atol=1e-12)
    assert np.all(out[-3:] == bias.astype(np.float64)[None, :])          # empty segment -> bias exactly
    assert np.all(np.isinf(smax[-3:])) and np.all(ssum[-3:] == 0)
    dOut = rng.standard_normal((T, H * C)).astype(np.float32)
    (o * torch.from_numpy(dOut).double()).sum().backward()
    dXL, dXR, datt, dbias = gat_edge_c.gat_edge_bwd(XL, XR, att, target, dOut, T, H, C)
    np.testing.assert_allclose(dXL, xl.grad[:E].numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dXR, xr.grad[E:].numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(datt, a.grad.reshape(-1).numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(dbias, dOut.astype(np.float64).sum(0), rtol=1e-12, atol=1e-12)


def test_synthetic_scene_invariants():
    idx, vals = gasfm_cpu.synthetic_observations(30, 900, 6000, seed=3)
    key = idx[0] * 900 + idx[1]
    assert np.all(np.diff(key) > 0)                                     # row-major sorted, no duplicates
    assert np.bincount(idx[1], minlength=900).min() >= 2
    assert np.bincount(idx[0], minlength=30).min() >= 8
    assert vals.shape == (idx.shape[1], 2) and vals.dtype == np.float32
    M = gasfm_cpu.dense_M_from_sparse(torch.from_numpy(idx), torch.from_numpy(vals), 30, 900)
    again = gasfm_cpu.make_scene(M, None)
    assert torch.equal(again["x"]["indices"], torch.from_numpy(idx))


LOSS_VARIANTS = ("shipped", "no_equalization", "normalize_all", "no_hinge")


@pytest.mark.parametrize("name", LOSS_VARIANTS)
def test_esfm_loss_oracle_matches_reference(name):
    """the dense loss restatement (incl. the gradient hook) against the reference's ESFMLoss."""
    import json
    g = load_golden("esfm_loss")
    lc = json.loads(str(g[f"{name}.conf_json"]))
    idx, vals = torch.from_numpy(g["indices"]), torch.from_numpy(g["values"])
    m, n = g["Ps"].shape[0], g["pts3D"].shape[1]
    Ps = torch.from_numpy(g["Ps"]).double().requires_grad_(True)
    X = torch.from_numpy(g["pts3D"]).double().requires_grad_(True)
    loss = gasfm_cpu.esfm_loss(Ps, X, idx, vals, m, n, 1e-4, lc["hinge_loss"], lc["hinge_loss_weight"],
                               lc["pts_grad_equalization_pre_perspective_divide"], lc["normalize_grad_wrt_valid_projections_only"])
    (loss * 3.0).backward()
    assert abs(loss.item() - float(g[f"{name}.f64.loss"])) < 1e-6       # observations are stored in fp32
    np.testing.assert_allclose(Ps.grad.numpy(), g[f"{name}.f64.dPs"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(X.grad.numpy(), g[f"{name}.f64.dpts3D"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["dpesfm_shipped_like", "dpesfm_skipconn"])
def test_dpesfm_oracle_matches_reference(name):
    import json
    g = load_golden(name)
    conf = json.loads(str(g["conf_json"]))["model"]
    params = {k[len("param."):]: torch.from_numpy(g[k]).double() for k in g.files if k.startswith("param.")}
    scene = gasfm_cpu.make_scene(torch.from_numpy(g["M"]).double(), torch.from_numpy(g["Ns"]).double())
    out = gasfm_cpu.set_of_set_forward(params, scene, conf["block_size"], conf["proj_feat_normalization"],
                                       conf["add_skipconn_for_residual_blocks"])
    for key in ("Ps_norm", "pts3D"):
        np.testing.assert_allclose(out[key].numpy(), g[f"out.f64.{key}"], rtol=1e-9, atol=1e-10)
