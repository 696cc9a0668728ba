"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Container-only (needs /root/reference):   python tests/golden/make_golden.py

Imports ``/root/reference/code`` through ``oracle/ref_import.py`` (third-party packages
that are absent offline are stubbed; ``torch_geometric.nn.GATv2Conv`` is the restatement
in ``oracle/gatv2conv.py``), runs the reference's own ``SceneData`` / ``GraphAttnSfMNet``
/ ``SetOfSetLayer`` / ``SparseMat`` on small seeded inputs, checks that
``oracle/gasfm_cpu.py`` reproduces them, and stores inputs, weights and outputs as
small ``.npz`` files.  The fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import gasfm_cpu, ref_import  # noqa: E402

MODEL_VARIANTS = {
    # name: (model-conf overrides, scene (m, n, density, seed))
    "tiny_shipped_like": (dict(), (10, 80, 0.35, 11)),
    "tiny_global2node_hidden1": (dict(global2view_and_global2scenepoint_enabled=True,
                                      n_hidden_layers_scenepoint_update=1, n_hidden_layers_view_update=1,
                                      n_hidden_layers_global_update=1, n_hidden_layers_proj_update=1,
                                      add_skipconn_from_init_projfeat=False), (9, 70, 0.4, 12)),
    "tiny_stateless_nonorm": (dict(stateful_global_features=False, use_norm_proj_update=False,
                                   n_feat_proj=12, n_heads=2), (12, 60, 0.3, 13)),
    "tiny_projective_depthhead": (dict(depth_head=dict(enabled=True, n_feat=20, n_hidden_layers=1),
                                       view_head=dict(enabled=True, n_hidden_layers=1,
                                                      normalize_output="Differentiable Chirality"),
                                       _calibrated=False), (8, 64, 0.4, 14)),
    "tiny_rot6d": (dict(view_head=dict(enabled=True, n_hidden_layers=2, rot_representation="6d"),
                        num_layers=2), (8, 48, 0.4, 15)),
}


def model_conf(**over):
    model = dict(type="graph_attn_sfm.GraphAttnSfMNet", n_heads=4, stateful_global_features=True,
                 global2view_and_global2scenepoint_enabled=False, n_feat_proj=16, n_feat_scenepoint=8,
                 n_feat_view=32, n_feat_global=64, num_layers=3,
                 n_hidden_layers_scenepoint_update=0, n_hidden_layers_view_update=0,
                 n_hidden_layers_global_update=0, n_hidden_layers_proj_update=0,
                 use_norm_proj_update=True, add_residual_skipconn_proj_update=True,
                 add_skipconn_from_init_projfeat=True, pos_emb_n_freq=0,
                 depth_head=dict(enabled=False, n_feat=128, n_hidden_layers=2),
                 view_head=dict(enabled=True, n_hidden_layers=2, rot_representation="quat"),
                 scenepoint_head=dict(enabled=True, n_hidden_layers=2))
    over = dict(over)
    calibrated = over.pop("_calibrated", True)
    model.update(over)
    return dict(dataset=dict(calibrated=calibrated), model=model)


def random_dense_scene(m, n, density, seed):
    """Dense M with deliberate degenerate tracks: an empty column, a single-view column."""
    g = torch.Generator().manual_seed(seed)
    mask = torch.rand(m, n, generator=g) < density
    mask[:, 3] = False                      # unobserved track
    mask[:, 5] = False
    mask[2, 5] = True                       # seen once -> dropped by MIN_N_VIEWS_PER_POINT
    xy = torch.randn(m, 2, n, generator=g) * 300 + 500
    M = (xy * mask[:, None, :]).reshape(2 * m, n)
    K = torch.eye(3).repeat(m, 1, 1)
    K[:, 0, 0] = K[:, 1, 1] = 800 + 50 * torch.rand(m, generator=g)
    K[:, 0, 2], K[:, 1, 2] = 500.0, 480.0
    Ns = torch.inverse(K)
    return M, Ns


def to_np(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def main():
    ref = ref_import.import_reference()
    torch.set_num_threads(4)

    # ---- index build + graphs (bit-exact ints) -------------------------------------------
    M, Ns = random_dense_scene(7, 50, 0.4, 3)
    M[2 * 6:, :] = 0
    M[2 * 6, :5] = 1.0                      # a view with < 8 points: excluded from view2global
    data = ref.SceneData.SceneData(M, Ns, torch.zeros(7, 3, 4), "golden", calibrated=True)
    fix = dict(M=M, Ns=Ns, values=data.x.values, indices=data.x.indices,
               cam_per_pts=data.x.cam_per_pts, pts_per_cam=data.x.pts_per_cam,
               shape=np.array(data.x.shape))
    for name, gw in data.graph_wrappers.items():
        fix[f"{name}.edge_index"] = gw.edge_index
        fix[f"{name}.valid_indices"] = gw.valid_indices
        fix[f"{name}.meta"] = np.array([gw.m, gw.n, gw.agg_dim, gw.n_agg_nodes])
    # the invariant documented (commented out) at SceneData.py:189-230
    nf = data.graph_wrappers["proj2view"].generate_node_features(data.x.to_torch_hybrid_sparse_coo())
    norm_M = ref.geo_utils.normalize_M(M, Ns)
    assert torch.equal(nf[: data.x.indices.shape[1]], norm_M[data.x.indices[0], data.x.indices[1], :])
    ora = gasfm_cpu.make_scene(M, Ns)
    assert torch.equal(ora["x"]["indices"], data.x.indices)
    assert torch.equal(ora["x"]["values"], data.x.values)
    assert torch.equal(ora["x"]["cam_per_pts"], data.x.cam_per_pts)
    assert torch.equal(ora["x"]["pts_per_cam"], data.x.pts_per_cam)
    for name, gw in data.graph_wrappers.items():
        assert torch.equal(ora["graphs"][name]["edge_index"], gw.edge_index), name
        assert torch.equal(ora["graphs"][name]["valid_indices"], gw.valid_indices), name
    np.savez_compressed(os.path.join(HERE, "index_build.npz"), **to_np(fix))
    print("index_build: E =", data.x.indices.shape[1])

    # ---- SparseMat pooling + DPESFM layer ------------------------------------------------
    torch.manual_seed(5)
    feat = torch.randn(data.x.indices.shape[1], 6)
    sm = ref.sparse_utils.SparseMat(feat, data.x.indices, data.x.cam_per_pts, data.x.pts_per_cam, (7, 50, 6))
    layer = ref.layers.SetOfSetLayer(6, 10)
    out = layer(sm)
    pfix = dict(indices=data.x.indices, feat=feat, shape=np.array([7, 50, 6]), sum0=sm.sum(0), sum1=sm.sum(1),
                mean0=sm.mean(0), mean1=sm.mean(1), sos_out=out.values)
    sd = {k: v for k, v in layer.state_dict().items()}
    pfix.update({f"param.{k}": v for k, v in sd.items()})
    assert torch.allclose(gasfm_cpu.sparse_sum(feat, data.x.indices, (7, 50, 6), 0), sm.sum(0))
    o2 = gasfm_cpu.set_of_set_layer(gasfm_cpu._P(sd), feat, data.x.indices, (7, 50, 6))
    assert torch.allclose(o2, out.values, atol=1e-6), (o2 - out.values).abs().max()
    np.savez_compressed(os.path.join(HERE, "pooling.npz"), **to_np(pfix))
    print("pooling ok")

    # ---- full model, three configurations -------------------------------------------------
    for name, (over, (m, n, dens, seed)) in MODEL_VARIANTS.items():
        conf_d = model_conf(**over)
        conf = ref.ConfigTree.from_dict(conf_d)
        torch.manual_seed(seed)
        model = ref.graph_attn_sfm.GraphAttnSfMNet(conf)
        # non-trivial norm parameters / conv biases so that every term is exercised
        with torch.no_grad():
            for k, v in model.named_parameters():
                if k.endswith("graph_conv.bias") or "norm" in k or k.endswith("2global.bias"):
                    v.add_(0.1 * torch.randn_like(v))
        M, Ns = random_dense_scene(m, n, dens, seed)
        data = ref.SceneData.SceneData(M, Ns, torch.zeros(m, 3, 4), name, calibrated=True)
        fix = dict(M=M, Ns=Ns)
        sd = model.state_dict()
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            mdl = model.to(dtype)
            d2 = ref.SceneData.SceneData(M.to(dtype), Ns.to(dtype), torch.zeros(m, 3, 4, dtype=dtype), name, calibrated=True)
            # the wrapper hard-codes float32 zeros for stateless queries (dataset_utils.py:571)
            mdl.zero_grad()
            if dtype == torch.float64:
                _patch_zero_query_dtype(ref, dtype)
            out = mdl(d2)
            w = torch.linspace(0.5, 1.5, out["Ps_norm"].numel(), dtype=dtype).reshape(out["Ps_norm"].shape)
            w2 = torch.linspace(-1.0, 1.0, out["pts3D"].numel(), dtype=dtype).reshape(out["pts3D"].shape)
            loss = (out["Ps_norm"] * w).sum() + (out["pts3D"] * w2).sum()
            if "depths" in out:
                loss = loss + (out["depths"].values ** 2).sum()
                fix[f"out.{tag}.depths"] = out["depths"].values
            loss.backward()
            fix[f"out.{tag}.Ps_norm"] = out["Ps_norm"]
            fix[f"out.{tag}.pts3D"] = out["pts3D"]
            fix[f"loss.{tag}"] = loss
            for k, v in mdl.named_parameters():
                fix[f"grad.{tag}.{k}"] = v.grad.detach().clone()   # clone: module.to() re-types .grad in place
            # oracle agreement (same dtype)
            params = {k: v.detach().clone().requires_grad_(True) for k, v in mdl.state_dict().items()}
            scene = gasfm_cpu.make_scene(M.to(dtype), Ns.to(dtype))
            mc = conf_d["model"]
            o = gasfm_cpu.gasfm_forward(params, scene, n_heads=mc["n_heads"], stateful=mc["stateful_global_features"],
                                        calibrated=conf_d["dataset"]["calibrated"],
                                        rot_representation=mc["view_head"].get("rot_representation", "quat"),
                                        normalize_output=mc["view_head"].get("normalize_output"))
            tol = 1e-5 if dtype == torch.float32 else 1e-12
            for key in ("Ps_norm", "pts3D"):
                err = (o[key] - out[key]).abs().max().item()
                assert err <= tol * max(1.0, out[key].abs().max().item()), (name, tag, key, err)
            l2 = (o["Ps_norm"] * w).sum() + (o["pts3D"] * w2).sum()
            if "depths" in o:
                assert (o["depths"] - out["depths"].values).abs().max().item() <= tol * 10
                l2 = l2 + (o["depths"] ** 2).sum()
            l2.backward()
            # gradients that are analytically zero (softmax shift-invariance) are pure rounding
            # noise, so errors are scaled by max(|g_param|, 1e-3 * largest gradient in the model)
            worst = 0.0
            gscale = 1e-3 * max(v.grad.abs().max().item() for v in mdl.parameters())
            for k, v in mdl.named_parameters():
                gerr = (params[k].grad - v.grad).abs().max().item() / max(gscale, v.grad.abs().max().item())
                worst = max(worst, gerr)
            assert worst < (2e-3 if dtype == torch.float32 else 1e-9), (name, tag, worst)
            print(f"{name} [{tag}]: E={scene['x']['indices'].shape[1]} oracle-vs-reference out ok, worst grad rel err {worst:.2e}")
        model.to(torch.float32)
        fix.update({f"param.{k}": v for k, v in sd.items()})
        fix["conf_json"] = np.array(__import__("json").dumps(conf_d))
        np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **to_np(fix))
    make_loss_and_dpesfm(ref)
    make_metric_and_sampling(ref)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


def make_metric_and_sampling(ref):
    """SURVEY 8(f2, f3): the reference's own per-step metric (evaluation.compute_core_errors) and view sub-sampling
    (SceneData.sample_data) on seeded inputs."""
    import importlib
    import types
    sys.modules.setdefault("PyCeres", types.ModuleType("PyCeres"))      # import-time only (utils/ba_functions.py)
    ev = importlib.import_module("evaluation")
    m, n = 9, 70
    M, Ns = random_dense_scene(m, n, 0.45, 31)
    y = torch.zeros(m, 3, 4)
    y[:, 0, 0] = torch.arange(m)                                        # lets the fixture record WHICH views were drawn
    data = ref.SceneData.SceneData(M, Ns, y, "metric", calibrated=True)
    g = torch.Generator().manual_seed(6)
    Ps = torch.randn(m, 3, 4, generator=g)
    Ps[:, 2, 3] += 3.0
    X = torch.cat((torch.randn(3, n, generator=g), 0.5 + torch.rand(1, n, generator=g)))     # not yet pflat'ed
    conf = ref.ConfigTree.from_dict({"dataset": {"calibrated": True}, "eval": {"calc_reprojerr_with_gtposes_for_depth_pred": False},
                                     "model": {"view_head": {"enabled": True}, "scenepoint_head": {"enabled": True}}})
    fix = dict(M=M, Ns=Ns, Ps_norm=Ps, pts3D=X)
    fix["our_repro.f32"] = np.float64(ev.compute_core_errors(data, {"Ps_norm": Ps, "pts3D": X}, conf)["our_repro"])
    data64 = ref.SceneData.SceneData(M.double(), Ns.double(), y.double(), "metric", calibrated=True)
    fix["our_repro.f64"] = np.float64(ev.compute_core_errors(data64, {"Ps_norm": Ps.double(), "pts3D": X.double()}, conf)["our_repro"])
    np.savez_compressed(os.path.join(HERE, "core_errors.npz"), **to_np(fix))
    print("core_errors ok:", float(fix["our_repro.f32"]), float(fix["our_repro.f64"]))

    fix = dict(M=M, Ns=Ns, y=y)
    for name, num_views, consecutive, seed in (("consecutive4", 4, True, 1), ("fraction", 0.6, True, 2), ("random5", 5, False, 3)):
        np.random.seed(seed)
        sub = ref.SceneData.sample_data(data, num_views, consecutive_views=consecutive)
        fix[f"{name}.args"] = np.array([num_views, float(consecutive), seed], dtype=np.float64)
        fix[f"{name}.view_ids"] = sub.y[:, 0, 0].to(torch.int64)
        fix[f"{name}.indices"], fix[f"{name}.values"] = sub.x.indices, sub.x.values
        fix[f"{name}.cam_per_pts"], fix[f"{name}.pts_per_cam"] = sub.x.cam_per_pts, sub.x.pts_per_cam
        fix[f"{name}.shape"] = np.array(sub.x.shape)
        fix[f"{name}.Ns"] = sub.Ns
        for key, w in sub.graph_wrappers.items():
            fix[f"{name}.graph.{key}"] = w.edge_index
    # greedy subset (SceneData.get_subset) and the rotational homography augmentation, same fixture file
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        sub = ref.SceneData.get_subset(data, 5)
    name = "subset5"
    fix[f"{name}.view_ids"] = sub.y[:, 0, 0].to(torch.int64)
    fix[f"{name}.indices"], fix[f"{name}.values"] = sub.x.indices, sub.x.values
    fix[f"{name}.cam_per_pts"], fix[f"{name}.pts_per_cam"] = sub.x.cam_per_pts, sub.x.pts_per_cam
    fix[f"{name}.shape"] = np.array(sub.x.shape)
    fix[f"{name}.Ns"] = sub.Ns
    for key, w in sub.graph_wrappers.items():
        fix[f"{name}.graph.{key}"] = w.edge_index
    g2 = torch.Generator().manual_seed(9)
    data_aug_src = ref.SceneData.SceneData(M, Ns, torch.randn(m, 3, 4, generator=g2), "aug", calibrated=True)
    fix["aug.y_in"] = data_aug_src.y
    for name, inplane, tilt, seed in (("aug_both", 15, 20, 4), ("aug_inplane", 30, None, 5), ("aug_tilt", None, 10, 6)):
        torch.manual_seed(seed)
        aug = ref.SceneData.apply_rotational_homography_aug(data_aug_src, inplane_rot_aug_max_angle=inplane, tilt_rot_aug_max_angle=tilt)
        fix[f"{name}.args"] = np.array([-1 if inplane is None else inplane, -1 if tilt is None else tilt, seed], dtype=np.float64)
        fix[f"{name}.indices"], fix[f"{name}.values"], fix[f"{name}.y"], fix[f"{name}.M"] = aug.x.indices, aug.x.values, aug.y, aug.M
    np.savez_compressed(os.path.join(HERE, "sample_data.npz"), **to_np(fix))
    print("sample_data ok:", {k: fix[k].tolist() for k in fix if k.endswith("view_ids")})


LOSS_VARIANTS = {
    "shipped": dict(hinge_loss=True, hinge_loss_weight=1, pts_grad_equalization_pre_perspective_divide=True,
                    normalize_grad_wrt_valid_projections_only=True),
    "no_equalization": dict(hinge_loss=True, hinge_loss_weight=0.5, pts_grad_equalization_pre_perspective_divide=False,
                            normalize_grad_wrt_valid_projections_only=False),
    "normalize_all": dict(hinge_loss=True, hinge_loss_weight=1, pts_grad_equalization_pre_perspective_divide=True,
                          normalize_grad_wrt_valid_projections_only=False),
    "no_hinge": dict(hinge_loss=False, hinge_loss_weight=1, pts_grad_equalization_pre_perspective_divide=True,
                     normalize_grad_wrt_valid_projections_only=True),
}


class _FakeCuda(torch.Tensor):
    """ESFMLoss asserts ``data.valid_pts.is_cuda`` (loss_functions.py:122); the fixtures are made on CPU."""
    @property
    def is_cuda(self):
        return True


def make_loss_and_dpesfm(ref):
    import importlib
    lf = importlib.import_module("loss_functions")
    # ---- ESFMLoss: value and gradients for four hook / hinge settings --------------------------------
    m, n = 9, 70
    M, Ns = random_dense_scene(m, n, 0.45, 21)
    data = ref.SceneData.SceneData(M, Ns, torch.zeros(m, 3, 4), "loss", calibrated=True)
    data.valid_pts = torch.Tensor._make_subclass(_FakeCuda, data.valid_pts)
    g = torch.Generator().manual_seed(5)
    Ps0 = torch.randn(m, 3, 4, generator=g)
    X0 = torch.cat((torch.randn(3, n, generator=g), torch.ones(1, n)))
    Ps0[:, 2, 3] += 2.0                      # most depths positive, some negative -> both branches
    fix = dict(M=M, Ns=Ns, Ps=Ps0, pts3D=X0, indices=data.x.indices, values=data.x.values)
    for name, lc in LOSS_VARIANTS.items():
        conf = ref.ConfigTree.from_dict({"model": {"view_head": {"enabled": True}, "scenepoint_head": {"enabled": True}},
                                         "loss": dict(infinity_pts_margin=1e-4, **lc)})
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            Ps = Ps0.detach().clone().to(dtype).requires_grad_(True)
            X = X0.detach().clone().to(dtype).requires_grad_(True)
            data._norm_M = data._norm_M.to(dtype)
            loss = lf.ESFMLoss(conf)({"Ps_norm": Ps, "pts3D": X}, data)
            (loss * 3.0).backward()
            fix[f"{name}.{tag}.loss"], fix[f"{name}.{tag}.dPs"], fix[f"{name}.{tag}.dpts3D"] = loss.detach(), Ps.grad.clone(), X.grad.clone()
            P2, X2 = Ps0.detach().clone().to(dtype).requires_grad_(True), X0.detach().clone().to(dtype).requires_grad_(True)
            l2 = gasfm_cpu.esfm_loss(P2, X2, data.x.indices, data.x.values, m, n, 1e-4, lc["hinge_loss"], lc["hinge_loss_weight"],
                                     lc["pts_grad_equalization_pre_perspective_divide"], lc["normalize_grad_wrt_valid_projections_only"])
            (l2 * 3.0).backward()
            tol = 1e-5 if dtype == torch.float32 else 1e-12
            assert abs(l2.item() - loss.item()) <= tol * max(1.0, abs(loss.item())), (name, tag)
            assert (P2.grad - Ps.grad).abs().max() <= tol * 10 and (X2.grad - X.grad).abs().max() <= tol * 10, (name, tag)
        fix[f"{name}.conf_json"] = np.array(__import__("json").dumps(lc))
    np.savez_compressed(os.path.join(HERE, "esfm_loss.npz"), **to_np(fix))
    print("esfm_loss ok:", {k: float(fix[f"{k}.f64.loss"]) for k in LOSS_VARIANTS})

    # ---- DPESFM SetOfSetNet ---------------------------------------------------------------------------
    for name, over in (("dpesfm_shipped_like", dict()), ("dpesfm_skipconn", dict(add_skipconn_for_residual_blocks=True, num_blocks=2, block_size=2))):
        model_d = dict(type="SetOfSet.SetOfSetNet", num_features=24, proj_feat_normalization=True,
                       add_skipconn_for_residual_blocks=False, num_blocks=1, block_size=3, pos_emb_n_freq=0,
                       depth_head=dict(enabled=False, n_feat=16, n_hidden_layers=1),
                       view_head=dict(enabled=True, n_hidden_layers=2, rot_representation="quat"),
                       scenepoint_head=dict(enabled=True, n_hidden_layers=2))
        model_d.update(over)
        conf_d = dict(dataset=dict(calibrated=True), model=model_d)
        torch.manual_seed(31)
        net = ref.SetOfSet.SetOfSetNet(ref.ConfigTree.from_dict(conf_d))
        m, n = 8, 60
        M, Ns = random_dense_scene(m, n, 0.4, 33)
        fix = dict(M=M, Ns=Ns, conf_json=np.array(__import__("json").dumps(conf_d)))
        fix.update({f"param.{k}": v.clone() for k, v in net.state_dict().items()})
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            mdl = net.to(dtype)
            mdl.zero_grad()
            d2 = ref.SceneData.SceneData(M.to(dtype), Ns.to(dtype), torch.zeros(m, 3, 4, dtype=dtype), name, calibrated=True)
            out = mdl(d2)
            w = torch.linspace(0.5, 1.5, out["Ps_norm"].numel(), dtype=dtype).reshape(out["Ps_norm"].shape)
            w2 = torch.linspace(-1.0, 1.0, out["pts3D"].numel(), dtype=dtype).reshape(out["pts3D"].shape)
            ((out["Ps_norm"] * w).sum() + (out["pts3D"] * w2).sum()).backward()
            fix[f"out.{tag}.Ps_norm"], fix[f"out.{tag}.pts3D"] = out["Ps_norm"].detach(), out["pts3D"].detach()
            for k, v in mdl.named_parameters():
                fix[f"grad.{tag}.{k}"] = v.grad.detach().clone()
            params = {k: v.detach().clone() for k, v in mdl.state_dict().items()}
            o = gasfm_cpu.set_of_set_forward(params, gasfm_cpu.make_scene(M.to(dtype), Ns.to(dtype)), model_d["block_size"],
                                             model_d["proj_feat_normalization"], model_d["add_skipconn_for_residual_blocks"])
            tol = 2e-5 if dtype == torch.float32 else 1e-11
            for key in ("Ps_norm", "pts3D"):
                assert (o[key] - out[key]).abs().max().item() <= tol * max(1.0, out[key].abs().max().item()), (name, tag, key)
        net.to(torch.float32)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **to_np(fix))
        print(name, "ok")


def _patch_zero_query_dtype(ref, dtype):
    """fp64 ground truth only: AxialAggregationGraphWrapper.generate_node_features creates
    float32 zeros for absent queries (dataset_utils.py:571); run the fp64 pass with the
    default dtype switched so that the concat is dtype-consistent."""
    orig = ref.dataset_utils.AxialAggregationGraphWrapper.generate_node_features
    if getattr(orig, "_patched", False):
        return

    def gen(self, M, x_agg=None):
        if x_agg is None:
            x_agg = torch.zeros((self.n_agg_nodes, M.shape[2]), dtype=M.dtype, device=self.device)
        return orig(self, M, x_agg)

    gen._patched = True
    ref.dataset_utils.AxialAggregationGraphWrapper.generate_node_features = gen


if __name__ == "__main__":
    main()
