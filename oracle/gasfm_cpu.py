"""CPU restatement of the GASFM graph-attention path (oracle, test-only; see oracle/__init__.py).

A functional re-expression of what the reference computes between
``SceneData.__init__`` and ``GraphAttnSfMNet.forward`` -- written against a flat
``{state_dict key: tensor}`` parameter dictionary instead of an ``nn.Module`` tree, so
the same weights can be fed to the reference (container only), to this oracle and to
the CUDA implementation.  Each function cites the reference lines it follows
(paths relative to ``/root/reference/code``).  Everything is plain torch on CPU,
materialising the same ``[E, .]`` intermediates the reference materialises, so it also
serves as the "port" CPU baseline that ``bench.py`` times.

Validated against the live reference by ``tests/golden/make_golden.py`` (container)
and against the committed fixtures by ``tests/test_oracle_golden.py`` (anywhere).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle.gatv2conv import gatv2_edge_softmax_aggregate

MIN_N_VIEWS_PER_POINT = 2   # utils/constants.py:2
MIN_N_POINTS_PER_VIEW = 8   # utils/constants.py:6


# ----------------------------------------------------------------------------------------
# a1: observation index (utils/dataset_utils.py:86-156, utils/geo_utils.py:689-703)
# ----------------------------------------------------------------------------------------
def valid_observation_mask(M):
    """[2m,n] measurement matrix -> [m,n] bool (dataset_utils.py:86-113).

    An observation is present iff |x|+|y| != 0; tracks seen in fewer than
    MIN_N_VIEWS_PER_POINT views are dropped entirely (column zeroed).
    """
    M = np.asarray(M)
    m = M.shape[0] // 2
    xy = M.reshape(m, 2, -1)
    valid = (np.abs(xy[:, 0, :]) + np.abs(xy[:, 1, :])) != 0
    valid[:, valid.sum(axis=0) < MIN_N_VIEWS_PER_POINT] = False
    return valid


def observation_index(M, Ns=None):
    """dense M[2m,n], Ns[m,3,3] -> dict(values[E,2] f32, indices[2,E] i64, cam_per_pts[n,1],
    pts_per_cam[m,1], shape) as ``M2sparse(M, normalize=True, Ns)`` returns
    (dataset_utils.py:116-156).  Indices are row-major (np.nonzero order)."""
    M = torch.as_tensor(M)
    m, n = M.shape[0] // 2, M.shape[1]
    valid = torch.from_numpy(valid_observation_mask(M.numpy()))
    cam_per_pts = valid.sum(dim=0).unsqueeze(1)
    pts_per_cam = valid.sum(dim=1).unsqueeze(1)
    indices = torch.from_numpy(np.array(np.nonzero(valid.numpy())))
    xy = M.reshape(m, 2, n)
    if Ns is not None:
        # geo_utils.normalize_M: (Ns @ [x; y; 1])[:2], zero where invalid.
        hom = torch.cat((xy, torch.ones(m, 1, n, dtype=M.dtype)), dim=1)
        xy = (torch.as_tensor(Ns) @ hom)[:, :2, :]
    values = xy.permute(0, 2, 1)[indices[0], indices[1], :]
    return dict(values=values.contiguous(), indices=indices, cam_per_pts=cam_per_pts,
                pts_per_cam=pts_per_cam, shape=(m, n, 2))


# ----------------------------------------------------------------------------------------
# a3/a4: axial aggregation graphs (dataset_utils.py:464-597, datasets/SceneData.py:153-239)
# ----------------------------------------------------------------------------------------
def axial_graph(m, n, agg_dim, valid_indices):
    """Edge list of AxialAggregationGraphWrapper (dataset_utils.py:511-537): element e is
    source node e; its target is aggregation node E + valid_indices[non_agg_dim][e]."""
    non_agg_dim = 1 - agg_dim
    n_agg = (m, n)[non_agg_dim]
    n_el = valid_indices.shape[1]
    edge_index = torch.stack((torch.arange(n_el, dtype=torch.int64), n_el + valid_indices[non_agg_dim]))
    return dict(m=m, n=n, agg_dim=agg_dim, non_agg_dim=non_agg_dim, n_agg_nodes=n_agg,
                valid_indices=valid_indices, edge_index=edge_index)


def scene_graphs(indices, m, n):
    """The four graphs of SceneData.create_axial_aggregation_graphs (SceneData.py:153-239)."""
    pts_per_view = torch.bincount(indices[0], minlength=m)
    views_per_pt = torch.bincount(indices[1], minlength=n)
    view_rows = torch.nonzero(pts_per_view >= MIN_N_POINTS_PER_VIEW)[:, 0]
    pt_cols = torch.nonzero(views_per_pt >= MIN_N_VIEWS_PER_POINT)[:, 0]
    v2g_idx = torch.stack((view_rows, torch.zeros_like(view_rows)))       # (m,1) matrix, SceneData.py:174-179
    p2g_idx = torch.stack((torch.zeros_like(pt_cols), pt_cols))           # (1,n) matrix, SceneData.py:182-187
    return {
        "proj2view": axial_graph(m, n, 1, indices),
        "proj2scenepoint": axial_graph(m, n, 0, indices),
        "view2global": axial_graph(m, 1, 0, v2g_idx),
        "scenepoint2global": axial_graph(1, n, 1, p2g_idx),
    }


def make_scene(M, Ns):
    """What the model reads from a SceneData: ``x`` and ``graph_wrappers``."""
    x = observation_index(M, Ns)
    m, n, _ = x["shape"]
    return dict(x=x, graphs=scene_graphs(x["indices"], m, n))


def scene_from_sparse(indices, values, m, n):
    """Same, from an already-sparse observation list (row-major sorted, deduplicated)."""
    indices = torch.as_tensor(indices, dtype=torch.int64)
    x = dict(values=torch.as_tensor(values), indices=indices,
             cam_per_pts=torch.bincount(indices[1], minlength=n).unsqueeze(1),
             pts_per_cam=torch.bincount(indices[0], minlength=m).unsqueeze(1), shape=(m, n, 2))
    return dict(x=x, graphs=scene_graphs(indices, m, n))


# ----------------------------------------------------------------------------------------
# a14: row / column mean pooling (utils/sparse_utils.py:406-419, :91-131)
# ----------------------------------------------------------------------------------------
def sparse_sum(values, indices, shape, dim):
    out_size = shape[0] if dim == 1 else shape[1]
    return values.new_zeros((out_size, values.shape[1])).index_add(0, indices[0 if dim == 1 else 1], values)


def sparse_mean(values, indices, shape, dim):
    """SparseMat.mean (sparse_utils.py:414-419): sum / count; empty rows or columns give 0/0 = nan
    there, exactly as the reference's division does."""
    cnt = torch.bincount(indices[0 if dim == 1 else 1], minlength=shape[0] if dim == 1 else shape[1])
    return sparse_sum(values, indices, shape, dim) / cnt.unsqueeze(1)


# ----------------------------------------------------------------------------------------
# parameter access helpers
# ----------------------------------------------------------------------------------------
class _P:
    """Prefix view on a flat state_dict."""

    def __init__(self, params, prefix=""):
        self.params, self.prefix = params, prefix

    def sub(self, name):
        return _P(self.params, f"{self.prefix}{name}.")

    def has(self, name):
        return f"{self.prefix}{name}" in self.params

    def __getitem__(self, name):
        return self.params[f"{self.prefix}{name}"]


def _linear(p, name, x):
    return F.linear(x, p[f"{name}.weight"], p[f"{name}.bias"] if p.has(f"{name}.bias") else None)


def _layer_norm(p, name, x):
    return F.layer_norm(x, (x.shape[-1],), p[f"{name}.weight"], p[f"{name}.bias"], 1e-5)


def _mlp(p, name, x):
    """get_linear_layers(feats, norm=False, no init/final activation) (models/layers.py:10-44):
    Linear (ReLU Linear)*, stored at indices 0, 2, 4, ..."""
    i = 0
    while p.has(f"{name}.{i}.weight"):
        if i > 0:
            x = F.relu(x)
        x = _linear(p, f"{name}.{i}", x)
        i += 2
    return x


# ----------------------------------------------------------------------------------------
# a5: GATv2 over an axial graph (layers.py:329-336 etc. + dataset_utils.py:539-590)
# ----------------------------------------------------------------------------------------
def gat_axial(p, graph, x_elements, x_agg, n_heads):
    """generate_node_features -> GATv2Conv -> extract_target_node_features.

    x_elements [E,d]; x_agg [T,d] or None (zeros, dataset_utils.py:569-571).  Returns [T, H*C].
    """
    n_el, t = x_elements.shape[0], graph["n_agg_nodes"]
    if x_agg is None:
        x_agg = x_elements.new_zeros((t, x_elements.shape[1]))
    nodes = torch.cat((x_elements, x_agg), dim=0)
    hc = p["lin_l.weight"].shape[0]
    x_l = _linear(p, "lin_l", nodes).view(-1, n_heads, hc // n_heads)
    x_r = _linear(p, "lin_r", nodes).view(-1, n_heads, hc // n_heads)
    out = gatv2_edge_softmax_aggregate(x_l, x_r, p["att"], graph["edge_index"]).reshape(-1, hc) + p["bias"]
    return out[n_el:]


def _norm_and_proj(p, name, x):
    """Sequential(LayerNorm, ReLU[, Linear]) (layers.py:295-303, 392-400, 497-505, 512-520)."""
    x = F.relu(_layer_norm(p, f"{name}.0", x))
    if p.has(f"{name}.2.weight"):
        x = _linear(p, f"{name}.2", x)
    return x


def _residual_update(p, proj_name, agg, prev):
    """Shared tail of Proj2View / Proj2ScenePoint / ViewAndScenePoint2Global
    (layers.py:341-357, 438-454, 583-599): project, + previous state, + mlp(relu(LN(.)))."""
    x = _linear(p, proj_name, agg) if p.has(f"{proj_name}.weight") else agg
    if prev is not None:
        x = prev + x
    return x + _mlp(p, "mlp", F.relu(_layer_norm(p, "norm_pre_mlp", x)))


def proj2view(p, graph, x_values, prev_view, n_heads):
    """Proj2View.forward (layers.py:321-361)."""
    q = None if prev_view is None else _norm_and_proj(p, "norm_and_proj_view2proj", prev_view)
    agg = gat_axial(p.sub("graph_conv"), graph, x_values, q, n_heads)
    return _residual_update(p, "proj_proj2view", agg, prev_view)


def proj2scenepoint(p, graph, x_values, prev_sp, n_heads):
    """Proj2ScenePoint.forward (layers.py:418-458)."""
    q = None if prev_sp is None else _norm_and_proj(p, "norm_and_proj_scenepoint2proj", prev_sp)
    agg = gat_axial(p.sub("graph_conv"), graph, x_values, q, n_heads)
    return _residual_update(p, "proj_proj2scenepoint", agg, prev_sp)


def view_and_scenepoint2global(p, g_v2g, g_p2g, view, sp, prev_global, n_heads):
    """ViewAndScenePoint2Global.forward (layers.py:538-603): only rows with >= 8 points and
    columns with >= 2 views are sources."""
    qv = None if prev_global is None else _norm_and_proj(p, "norm_and_proj_global2view", prev_global)
    qp = None if prev_global is None else _norm_and_proj(p, "norm_and_proj_global2scenepoint", prev_global)
    v = gat_axial(p.sub("graph_conv_view2global"), g_v2g, view[g_v2g["valid_indices"][0]], qv, n_heads)
    s = gat_axial(p.sub("graph_conv_scenepoint2global"), g_p2g, sp[g_p2g["valid_indices"][1]], qp, n_heads)
    return _residual_update(p, "proj_view_and_scenepoint2global", torch.cat((v, s), dim=1), prev_global)


def global_to_node(p, node_name, glob, prev):
    """Global2View / Global2ScenePoint (layers.py:634-662, 693-721)."""
    x = _linear(p, f"lin_{node_name}", F.relu(_layer_norm(p, f"{node_name}_norm_layer", prev)))
    x = x + _linear(p, "lin_global", F.relu(_layer_norm(p, "global_norm_layer", glob)))
    if p.has("mlp.0.weight"):
        x = _mlp(p, "mlp", F.relu(x))
    return prev + x


def global_feature_update(p, graphs, x_values, prev_sp, prev_view, prev_global, n_heads, output_global):
    """GraphAttnSfMGlobalFeatureUpdate.forward (layers.py:810-870)."""
    sp = proj2scenepoint(p.sub("proj2scenepoint"), graphs["proj2scenepoint"], x_values, prev_sp, n_heads)
    view = proj2view(p.sub("proj2view"), graphs["proj2view"], x_values, prev_view, n_heads)
    glob = None
    if p.has("view_and_scenepoint2global.mlp.0.weight"):
        glob = view_and_scenepoint2global(p.sub("view_and_scenepoint2global"), graphs["view2global"],
                                          graphs["scenepoint2global"], view, sp, prev_global, n_heads)
    if p.has("global2view.lin_view.weight"):
        sp = global_to_node(p.sub("global2scenepoint"), "scenepoint", glob, sp)
        view = global_to_node(p.sub("global2view"), "view", glob, view)
    return (sp, view, glob) if output_global else (sp, view)


def projection_feature_update(p, indices, x_values, sp, view, glob):
    """GraphAttnSfMProjectionFeatureUpdate.forward (layers.py:911-956)."""
    sp = _linear(p, "lin_scenepoint", F.relu(_layer_norm(p, "scenepoint_norm_layer", sp)))
    view = _linear(p, "lin_view", F.relu(_layer_norm(p, "view_norm_layer", view)))
    glob = _linear(p, "lin_global", F.relu(_layer_norm(p, "global_norm_layer", glob)))
    new = (_linear(p, "lin_proj", x_values) + sp[indices[1]] + view[indices[0]] + glob) / 4
    if p.has("mlp.0.weight"):
        new = _mlp(p, "mlp", F.relu(new))
    return new


def gasfm_layer(p, scene, x_raw, x0, prev_sp, prev_view, prev_global, n_heads):
    """GraphAttnSfMLayer.forward (layers.py:222-263).

    Quirk kept from the reference: ``relu_on_projection_features`` is an in-place ReLU
    (layers.py:982-984).  With ``use_norm_proj_update`` (all shipped confs) it acts on the
    fresh LayerNorm output; without it, it overwrites the layer's *input* values, so the
    residual branch sees relu(x_raw) as well (and block 0 rectifies the embedding that
    later blocks concatenate -- handled in ``gasfm_forward``)."""
    if p.has("prev_projfeat_norm_layer.weight"):
        x = F.relu(_layer_norm(p, "prev_projfeat_norm_layer", x_raw))
    else:
        x = F.relu(x_raw)
        x_raw = x
    sp, view, glob = global_feature_update(p.sub("global_feature_update"), scene["graphs"], x,
                                           prev_sp, prev_view, prev_global, n_heads, True)
    pf = p.sub("projection_feature_update")
    d_in = pf["lin_proj.weight"].shape[1]
    x_cat = torch.cat((x, x0), dim=1) if d_in != x.shape[1] else x      # layers.py:245-251
    new = projection_feature_update(pf, scene["x"]["indices"], x_cat, sp, view, glob)
    skip = x_raw
    if p.has("skip_projection.lin_proj.weight"):                         # layers.py:254-261
        if p.has("residual_skipconn_proj_norm_layer.weight"):
            skip = F.relu(_layer_norm(p, "residual_skipconn_proj_norm_layer", skip))
        skip = _linear(p, "skip_projection.lin_proj", skip)
    return skip + new, sp, view, glob


def quaternion_to_matrix(q):
    """pytorch3d.transforms.quaternion_to_matrix semantics (real part first, 2/|q|^2 scale);
    used at models/baseNet.py:48."""
    r, i, j, k = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    rows = (1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
            s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
            s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j))
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def rotation_6d_to_matrix(d6):
    """pytorch3d.transforms.rotation_6d_to_matrix semantics (Gram-Schmidt on two 3-vectors,
    result rows b1, b2, b1 x b2); used at models/baseNet.py:43."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def decode_views(m_out, calibrated=True, rot_representation="quat", normalize_output=None):
    """BaseNet.extract_view_outputs (models/baseNet.py:38-85)."""
    if calibrated:
        if rot_representation == "quat":
            rot = quaternion_to_matrix(m_out[:, :4])
        elif rot_representation == "6d":
            rot = rotation_6d_to_matrix(m_out[:, :6])
        elif rot_representation == "svd":                                # utils/geo_utils.py:25-31
            u, _, v = torch.svd(m_out[:, :9].reshape(-1, 3, 3))
            vt = v.transpose(1, 2)
            det = torch.det(u @ vt).view(-1, 1, 1)
            rot = u @ torch.cat((vt[:, :2, :], vt[:, -1:, :] * det), 1)
        else:
            raise ValueError(rot_representation)
        return torch.cat((rot, m_out[:, -3:].unsqueeze(-1)), dim=-1)
    Ps = m_out.reshape(-1, 3, 4)
    if normalize_output == "Chirality":
        Ps = Ps * (torch.sign(Ps[:, :3, :3].det()) / Ps[:, 2, :3].norm(dim=1)).reshape(-1, 1, 1)
    elif normalize_output == "Differentiable Chirality":
        Ps = Ps * (F.softsign(Ps[:, :3, :3].det() * 10e3) / Ps[:, 2, :3].norm(dim=1)).reshape(-1, 1, 1)
    elif normalize_output == "Frobenius":
        Ps = Ps / Ps.norm(dim=(1, 2), p="fro", keepdim=True)
    return Ps


def gasfm_forward(params, scene, n_heads=4, stateful=True, calibrated=True, rot_representation="quat",
                  normalize_output=None, return_features=False):
    """GraphAttnSfMNet.forward (models/graph_attn_sfm.py:117-185, models/baseNet.py:34-92)."""
    p = _P(params)
    values = scene["x"]["values"].to(params["embed.post_embed_lin.weight"].dtype)
    x0 = _linear(p, "embed.post_embed_lin", values)                      # layers.py:1009-1015
    if not p.has("equivariant_blocks.0.prev_projfeat_norm_layer.weight"):
        x0 = F.relu(x0)          # in-place ReLU aliasing in block 0, see gasfm_layer
    x, sp, view, glob = x0, None, None, None
    i = 0
    while p.has(f"equivariant_blocks.{i}.projection_feature_update.lin_proj.weight"):
        x, sp, view, glob = gasfm_layer(p.sub(f"equivariant_blocks.{i}"), scene, x, x0,
                                        sp if stateful else None, view if stateful else None,
                                        glob if stateful else None, n_heads)
        i += 1
    sp_f, view_f = global_feature_update(p.sub("final_global_update"), scene["graphs"], x,
                                         sp if stateful else None, view if stateful else None,
                                         glob if stateful else None, n_heads, False)
    m_out = _mlp(p, "view_head", F.relu(view_f))
    n_out = _mlp(p, "scenepoint_head", F.relu(sp_f)).T
    out = {
        "Ps_norm": decode_views(m_out, calibrated, rot_representation, normalize_output),
        "pts3D": torch.cat((n_out, n_out.new_ones(1, n_out.shape[1])), dim=0),
    }
    if p.has("depth_head.0.weight"):                                     # graph_attn_sfm.py:153-162
        out["depths"] = _mlp(p, "depth_head", x)
    if return_features:
        out.update(proj_features=x, scenepoint_features=sp_f, view_features=view_f, global_features=glob)
    return out


# ----------------------------------------------------------------------------------------
# a14: DPESFM set-of-set layer (layers.py:100-147) on the pooling primitive
# ----------------------------------------------------------------------------------------
def set_of_set_layer(p, x_values, indices, shape):
    """SetOfSetLayer.forward: means over columns / rows / everything, linear maps, average of 4."""
    pg = p.sub("global_feature_update")
    sp = _linear(pg, "lin_scenepoint", _nan_to_zero_mean(x_values, indices, shape, 0))
    view = _linear(pg, "lin_view", _nan_to_zero_mean(x_values, indices, shape, 1))
    glob = _linear(pg, "lin_global", x_values.mean(dim=0, keepdim=True))
    return (_linear(p.sub("projection_feature_update"), "lin_proj", x_values) + sp[indices[1]] + view[indices[0]] + glob) / 4


def _nan_to_zero_mean(values, indices, shape, dim):
    """sparse_utils.sparse_mean(...).to_dense() (sparse_utils.py:91-131): empty rows/cols are
    unspecified in the sparse result and become 0 when densified."""
    cnt = torch.bincount(indices[0 if dim == 1 else 1], minlength=shape[0] if dim == 1 else shape[1]).unsqueeze(1)
    return sparse_sum(values, indices, shape, dim) / cnt.clamp(min=1)


# ----------------------------------------------------------------------------------------
# f1: ESFM reprojection loss, dense like the reference (loss_functions.py:69-123)
# ----------------------------------------------------------------------------------------
def esfm_loss(Ps, pts3D, indices, values, m, n, margin=1e-4, hinge_loss=True, hinge_loss_weight=1.0,
              grad_equalization=True, normalize_valid_only=True):
    """ESFMLoss.forward on dense [m,3,n] tensors, including the gradient hook on ``Ps @ pts3D``
    (loss_functions.py:101-110).  ``indices/values``: the scene's observations (normalised)."""
    valid = torch.zeros(m, n, dtype=torch.bool)
    valid[indices[0], indices[1]] = True
    norm_M = torch.zeros(m, 2, n, dtype=Ps.dtype)
    norm_M[indices[0], :, indices[1]] = values.to(Ps.dtype)
    pts_2d = Ps @ pts3D
    z = pts_2d[:, 2, :]
    ok = (z >= margin) if hinge_loss else (z.abs() >= margin)           # geo_utils.py:721-726
    if not hinge_loss:
        hinge_loss_weight = 0
    if grad_equalization and pts_2d.requires_grad:
        if normalize_valid_only:
            count = max(1, int((valid & ok).sum()))
            pts_2d.register_hook(lambda g: torch.where(ok[:, None, :].expand(-1, 3, -1), F.normalize(g, dim=1) / count, g))
        else:
            pts_2d.register_hook(lambda g: F.normalize(g, dim=1) / valid.sum())
    hinge = (margin - z) * hinge_loss_weight
    proj = pts_2d / torch.where(ok, z, torch.ones_like(z)).unsqueeze(1)
    err = (proj[:, 0:2, :] - norm_M).norm(dim=1)
    return torch.where(ok, err, hinge)[valid].mean()


# ----------------------------------------------------------------------------------------
# f4: DPESFM SetOfSetNet (models/SetOfSet.py) on the set-of-set layer above
# ----------------------------------------------------------------------------------------
def set_of_set_forward(params, scene, block_size, proj_feat_normalization, add_skipconn):
    """SetOfSetNet.forward (models/SetOfSet.py:102-142) for calibrated / quaternion outputs."""
    p = _P(params)
    x = scene["x"]["values"].to(next(iter(params.values())).dtype)
    idx, shape = scene["x"]["indices"], scene["x"]["shape"]

    def normalize(v):                                                    # layers.py:977 (no norm layer)
        return v - v.mean(dim=0, keepdim=True)

    b = 0
    while p.has(f"equivariant_blocks.{b}.layers.0.projection_feature_update.lin_proj.weight"):
        pb = p.sub(f"equivariant_blocks.{b}")
        xl = x
        for i in range(block_size):
            xl = set_of_set_layer(pb.sub(f"layers.{i}"), xl, idx, shape)
            if i < block_size - 1:
                if proj_feat_normalization:
                    xl = normalize(xl)
                xl = F.relu(xl)
        if add_skipconn:
            skip = x
            if pb.has("skip_projection.lin_proj.weight"):
                skip = _linear(pb, "skip_projection.lin_proj", skip)
                if proj_feat_normalization:
                    skip = normalize(skip)
            xl = skip + xl
        x = F.relu(xl)
        b += 1
    pg = p.sub("final_global_update")
    n_in = _linear(pg, "lin_scenepoint", _nan_to_zero_mean(x, idx, shape, 0))
    m_in = _linear(pg, "lin_view", _nan_to_zero_mean(x, idx, shape, 1))
    m_out = _mlp(p, "view_head", F.relu(m_in))
    n_out = _mlp(p, "scenepoint_head", F.relu(n_in)).T
    return {"Ps_norm": decode_views(m_out), "pts3D": torch.cat((n_out, n_out.new_ones(1, n_out.shape[1])), dim=0)}


# ----------------------------------------------------------------------------------------
# synthetic scenes -- sparse-first, never builds the dense M
# ----------------------------------------------------------------------------------------
from gasfm_b200.synthetic import synthetic_observations  # noqa: E402,F401  (input generator shared with bench.py; no path arithmetic)


def synthetic_scene(m, n, n_obs, seed, banded=True):
    idx, vals = synthetic_observations(m, n, n_obs, seed, banded)
    return scene_from_sparse(torch.from_numpy(idx), torch.from_numpy(vals), m, n)


def dense_M_from_sparse(indices, values, m, n):
    """Dense [2m,n] measurement matrix with the given observations (identity Ns)."""
    M = torch.zeros(m, 2, n, dtype=values.dtype)
    M[indices[0], :, indices[1]] = values
    return M.reshape(2 * m, n)
