"""Mirror of the reference's ``code/models/layers.py``: the same module classes, constructor
signatures, ``forward`` signatures and parameter names (so ``gasfm_*.pt`` state_dicts load
unchanged), with the per-observation work running on the sm_100a kernels of ``gasfm_b200.ops``.

What changes under the hood, compared with the reference implementation:
  * the projection features stay a ``SparseMat`` everywhere; no ``sparse_coo_tensor().coalesce()``
    round trips (reference: layers.py:823,922,545-549,561-565) and no per-call index asserts;
  * GATv2 aggregation = ``lin_l`` on the E observation rows only + one fused kernel
    (scores, segment softmax, weighted sum) over the scene's cached CSR/CSC segments;
  * LayerNorm+ReLU on observation features is one kernel; the observation update
    ``(lin_proj(x) + sp[col] + view[row] + glob)/4 + skip`` (layers.py:941-945, 254-261) is one GEMM
    plus one gather-add kernel, without materialising the ``cat(x, x0)`` (layers.py:245-251).
Node-level work (``[m, .]``, ``[n, .]``, ``[1, .]`` LayerNorms, Linears, MLPs) is small and stays on
cuBLAS through torch.
"""
import torch
from torch import nn
from torch.nn import ReLU, Sequential, Module, Identity
from torch.nn import functional as F

from .. import _lib, ops
from ..index import SegmentPlan, index_for, one_segment_ptr, single_segment_chunk
from ..utils.sparse_utils import SparseMat
from ..utils import sparse_utils
from ..utils.pos_enc_utils import get_embedder
from .gatv2 import GATv2Conv


class Linear(nn.Linear):
    """``nn.Linear`` (same parameters / state_dict) whose forward runs the tcgen05 3xTF32 GEMM when the
    input has thousands of rows (observation- and point-level features) and cuBLAS fp32 otherwise."""

    def forward(self, x):
        if x.dim() == 2 and x.is_cuda and x.dtype == torch.float32:
            return ops.linear(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    """``nn.LayerNorm`` (same parameters); ``relu_(ln(x))`` patterns call ``ln_relu`` to get one fused kernel."""

    def ln_relu(self, x):
        if (x.dim() == 2 and x.is_cuda and x.dtype == torch.float32 and x.shape[0] >= 256
                and ops.ln_relu_width_supported(x.shape[1]) and self.elementwise_affine and self.bias is not None):
            return ops.ln_relu(x, self.weight, self.bias, self.eps)
        return F.relu(F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps))


class _NormReluProj(Sequential):
    """Sequential(LayerNorm, ReLU[, Linear]) with the state_dict layout of the reference's
    ``norm_and_proj_*`` modules (indices 0 and 2), evaluated with the fused LayerNorm+ReLU kernel."""

    def forward(self, x):
        x = self[0].ln_relu(x)
        return self[2](x) if len(self) > 2 else x


def get_linear_layers(feats, init_activation=False, final_activation=False, norm=True):
    """[LN] ReLU? (Linear [LN] ReLU)* Linear ([LN] ReLU)? -- module indices as in layers.py:10-44."""
    assert len(feats) >= 2
    layers = []

    def act(width):
        if norm:
            layers.append(LayerNorm(width))
        layers.append(ReLU(inplace=True))

    if init_activation:
        act(feats[0])
    for a, b in zip(feats[:-2], feats[1:-1]):
        layers.append(Linear(a, b))
        act(b)
    layers.append(Linear(feats[-2], feats[-1]))
    if final_activation:
        act(feats[-1])
    return Sequential(*layers)


def _round_up(width, heads):
    return width if width % heads == 0 else width + heads - width % heads


class Parameter3DPts(torch.nn.Module):
    def __init__(self, n_pts):
        super().__init__()
        self.pts_3d = torch.nn.Parameter(torch.normal(mean=0, std=0.1, size=(3, n_pts)))

    def forward(self):
        return self.pts_3d


# ---------------------------------------------------------------------------------------------
# plans for the reference's graph wrappers
# ---------------------------------------------------------------------------------------------
_PLAN_ATTR = "_gasfm_b200_plan"


def plan_for(graph_wrapper, proj_features=None):
    """Segment plan of an ``AxialAggregationGraphWrapper`` (ours or the reference's, duck-typed).

    Row / column aggregation over the observation matrix maps to the scene's cached CSR (views)
    or CSC (tracks); the single-target global graphs get their own one-segment plan whose ``perm``
    lists the valid source rows (views with >= 8 points / tracks with >= 2 views)."""
    if graph_wrapper.n_agg_nodes == 1 and (graph_wrapper.m == 1 or graph_wrapper.n == 1):
        ids = graph_wrapper.valid_indices[graph_wrapper.agg_dim]
        plan = getattr(graph_wrapper, _PLAN_ATTR, None)
        if plan is None or plan.seg_ptr.device != ids.device or plan.n_edges != ids.numel():
            dev = ids.device
            k = int(ids.numel())
            n_src = graph_wrapper.m if graph_wrapper.agg_dim == 0 else graph_wrapper.n
            seg_ptr = one_segment_ptr(k, dev)
            perm = None if k == n_src else ids.to(torch.int32).contiguous()
            with _lib.device_guard(dev):
                plan = SegmentPlan(seg_ptr, perm, 1, k, single_segment_chunk(k), dev)
            setattr(graph_wrapper, _PLAN_ATTR, plan)
        plan.shard = getattr(graph_wrapper, "shard", None)     # set for track-sharded scenes (gasfm_b200.dist)
        return plan
    assert proj_features is not None
    idx = index_for(proj_features)
    if graph_wrapper.agg_dim == 1:
        idx.by_view.shard = getattr(graph_wrapper, "shard", None)
        return idx.by_view
    return idx.by_track


# ---------------------------------------------------------------------------------------------
# DPESFM set-of-set layers (row / column mean pooling), layers.py:87-147
# ---------------------------------------------------------------------------------------------
class SetOfSetLayer(Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.global_feature_update = SetOfSetGlobalFeatureUpdate(d_in, d_out)
        self.projection_feature_update = SetOfSetProjectionFeatureUpdate(d_in, d_out)

    def forward(self, x):
        scenepoint_features, view_features, global_features = self.global_feature_update(x)
        return self.projection_feature_update(scenepoint_features, view_features, global_features, x)


class SetOfSetGlobalFeatureUpdate(Module):
    def __init__(self, d_in, d_out, output_global=True):
        super().__init__()
        self.lin_scenepoint = Linear(d_in, d_out)
        self.lin_view = Linear(d_in, d_out)
        self.output_global = output_global
        if output_global:
            self.lin_global = Linear(d_in, d_out)

    def forward(self, x):
        scenepoint_features = self.lin_scenepoint(sparse_utils.sparse_mean(x, dim=0))   # [n,d_in] -> [n,d_out]
        view_features = self.lin_view(sparse_utils.sparse_mean(x, dim=1))               # [m,d_in] -> [m,d_out]
        if not self.output_global:
            return scenepoint_features, view_features
        global_features = self.lin_global(sparse_utils.sparse_mean(x, dim=(0, 1))[None, :])
        return scenepoint_features, view_features, global_features


class SetOfSetProjectionFeatureUpdate(Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.lin_proj = Linear(d_in, d_out)

    def forward(self, scenepoint_features, view_features, global_features, x):
        proj = ops.linear(x.values, 0.25 * self.lin_proj.weight, 0.25 * self.lin_proj.bias)
        new = ops.edge_update(proj, None, None, scenepoint_features, view_features, global_features, None,
                              index_for(x), 1.0, 0.25)
        return x.with_values(new)


# ---------------------------------------------------------------------------------------------
# GASFM layers
# ---------------------------------------------------------------------------------------------
class _LazyFeatureCat(SparseMat):
    """``cat(x, x0)`` along the feature axis that is only materialised if somebody reads
    ``.values``; the projection update consumes the two parts separately."""

    def __init__(self, main, extra):
        self._main, self._extra = main, extra
        self._values = None
        self.indices, self.cam_per_pts, self.pts_per_cam = main.indices, main.cam_per_pts, main.pts_per_cam
        self.shape = (main.shape[0], main.shape[1], main.shape[2] + extra.shape[2])
        self.device = main.device
        idx = getattr(main, "_gasfm_b200_index", None)
        if idx is not None:
            self._gasfm_b200_index = idx

    @property
    def values(self):
        if self._values is None:
            self._values = torch.cat((self._main.values, self._extra.values), dim=1)
        return self._values

    @values.setter
    def values(self, v):
        self._values = v


class GraphAttnSfMLayer(Module):
    def __init__(
        self,
        n_feat_proj_in,
        n_feat_proj_out,
        n_feat_scenepoint_hidden,
        n_feat_view_hidden,
        n_feat_global_hidden,
        n_feat_proj2scenepoint_agg = None,
        n_feat_proj2view_agg = None,
        n_feat_scenepoint2global_agg = None,
        n_feat_view2global_agg = None,
        use_norm_proj_update = True,
        add_residual_skipconn_proj_update = True,
        n_feat_skipconn_init_projfeat_in = None,
        n_heads = 1,
        stateful = True,
        global2view_and_global2scenepoint_enabled = True,
        n_hidden_layers_scenepoint_update = 0,
        n_hidden_layers_view_update = 0,
        n_hidden_layers_global_update = 0,
        n_hidden_layers_proj_update = 0,
    ):
        super().__init__()
        self.use_norm_proj_update = use_norm_proj_update
        self.add_residual_skipconn_proj_update = add_residual_skipconn_proj_update
        self.add_skipconn_from_init_projfeat = n_feat_skipconn_init_projfeat_in is not None
        self.n_feat_skipconn_init_projfeat_in = n_feat_skipconn_init_projfeat_in or 0

        if self.use_norm_proj_update:
            self.prev_projfeat_norm_layer = LayerNorm(n_feat_proj_in)
        self.global_feature_update = GraphAttnSfMGlobalFeatureUpdate(
            n_feat_proj_in, n_feat_scenepoint_hidden, n_feat_view_hidden,
            n_feat_proj2scenepoint_agg = n_feat_proj2scenepoint_agg,
            n_feat_proj2view_agg = n_feat_proj2view_agg,
            n_feat_global_out = n_feat_global_hidden,
            n_feat_scenepoint2global_agg = n_feat_scenepoint2global_agg,
            n_feat_view2global_agg = n_feat_view2global_agg,
            output_global = True,
            n_heads = n_heads,
            stateful = stateful,
            global2view_and_global2scenepoint_enabled = global2view_and_global2scenepoint_enabled,
            n_hidden_layers_scenepoint_update = n_hidden_layers_scenepoint_update,
            n_hidden_layers_view_update = n_hidden_layers_view_update,
            n_hidden_layers_global_update = n_hidden_layers_global_update,
        )
        self.projection_feature_update = GraphAttnSfMProjectionFeatureUpdate(
            n_feat_proj_in + self.n_feat_skipconn_init_projfeat_in,
            n_feat_scenepoint_hidden, n_feat_view_hidden, n_feat_global_hidden, n_feat_proj_out,
            n_hidden_layers_proj_update = n_hidden_layers_proj_update,
            normalize_global_features = True,
        )
        if self.add_residual_skipconn_proj_update:
            if n_feat_proj_in == n_feat_proj_out:
                self.skip_projection = None
            else:
                if self.use_norm_proj_update:
                    self.residual_skipconn_proj_norm_layer = LayerNorm(n_feat_proj_in)
                self.skip_projection = ProjLayer(n_feat_proj_in, n_feat_proj_out)

    def forward(
        self,
        prev_projection_features,
        graph_structure,
        prev_scenepoint_features = None,
        prev_view_features = None,
        prev_global_features = None,
        skipconn_init_projfeat = None,
    ):
        raw = prev_projection_features
        norm = self.prev_projfeat_norm_layer if self.use_norm_proj_update else None
        plain_residual = self.add_residual_skipconn_proj_update and self.skip_projection is None and norm is not None
        gfu, pfu = self.global_feature_update, self.projection_feature_update
        d_main = raw.shape[2]
        conv_sp, conv_v = gfu.proj2scenepoint.graph_conv, gfu.proj2view.graph_conv
        # lin_l of both attention graphs and lin_proj all read x = relu(LN(x_raw))
        projections = [(conv_sp.lin_l.weight, conv_sp.lin_l.bias), (conv_v.lin_l.weight, conv_v.lin_l.bias),
                       (0.25 * pfu.lin_proj.weight[:, :d_main], 0.25 * pfu.lin_proj.bias)]
        if plain_residual and ops.edge_block_supported(raw.values, norm.weight, projections):
            # the whole observation-level front end of the block as ONE autograd node: LN+ReLU, the grouped projection
            # GEMM, and in backward the concatenated input gradient + the LN backward with the residual gradient;
            # under activation recompute it keeps only x_raw and the LayerNorm statistics (ops.EdgeBlockContext)
            rc = ops.EdgeBlockContext(ops.activation_recompute_enabled())
            xl_sp, xl_v, proj, raw_vals = ops.edge_block_project(raw.values, norm.weight, norm.bias, norm.eps, rc, projections)
            x, raw = raw.with_values(None, n_feat=d_main), raw.with_values(raw_vals)
            # (tensor, rebuild-in-backward getter, slot where the attention backward leaves the row maxima of dXL)
            xl_sp, xl_v = (xl_sp, rc.xl_getter(0), rc.slot(0)), (xl_v, rc.xl_getter(1), rc.slot(1))
            update_slot = rc.slot(2)
        else:
            update_slot = None
            if plain_residual and ops.ln_relu_width_supported(d_main):
                # x_raw feeds LN+ReLU and the residual: one autograd node, so that the two gradients are summed
                # inside the LN+ReLU backward kernel (ops.ln_relu_with_skip)
                y, raw_vals = ops.ln_relu_with_skip(raw.values, norm.weight, norm.bias, norm.eps)
                x, raw = raw.with_values(y), raw.with_values(raw_vals)
            else:
                x = relu_on_projection_features(None, _fused_norm=(raw, norm))
            if norm is None:
                # The reference's ReLU is in-place (layers.py:982-984): without a norm layer in front it
                # also rectifies the values the residual branch reads.
                raw = x
            # one autograd node, so that the three input gradients are summed inside the GEMM epilogues (ops.linear_multi)
            xl_sp, xl_v, proj = ops.linear_multi(x.values, projections)
        scenepoint_features, view_features, global_features = gfu(
            x, graph_structure,
            prev_scenepoint_features = prev_scenepoint_features,
            prev_view_features = prev_view_features,
            prev_global_features = prev_global_features,
            projected_sources = (xl_sp, xl_v),
        )
        feats = x
        if self.add_skipconn_from_init_projfeat:
            assert skipconn_init_projfeat is not None
            assert skipconn_init_projfeat.values.shape[1] == self.n_feat_skipconn_init_projfeat_in
            feats = _LazyFeatureCat(x, skipconn_init_projfeat)

        residual = None
        if self.add_residual_skipconn_proj_update:
            residual = raw
            if self.skip_projection is not None:
                if self.use_norm_proj_update:
                    residual = relu_on_projection_features(None, _fused_norm=(residual, self.residual_skipconn_proj_norm_layer))
                residual = self.skip_projection(residual)
        projection_features = pfu(scenepoint_features, view_features, global_features, feats, residual = residual,
                                  projected = proj, rowmax_slot = update_slot)
        return projection_features, scenepoint_features, view_features, global_features


class _AxialAttentionUpdate(Module):
    """Shared body of Proj2View / Proj2ScenePoint: query = Linear(ReLU(LN(prev))), GATv2
    aggregation over rows / columns, projection, residual, pre-norm MLP (layers.py:321-361, 418-458)."""

    def _init_axial(self, n_feat_proj_in, n_feat_out, n_heads, stateful, use_norm_pre_mlp, n_feat_agg, n_hidden):
        self.n_feat_proj_in = n_feat_proj_in
        self.stateful = stateful
        self.use_norm_pre_mlp = use_norm_pre_mlp
        if n_feat_agg is None:
            n_feat_agg = _round_up(n_feat_proj_in, n_heads)
        assert n_feat_agg % n_heads == 0
        query_proj = None
        if stateful:
            mods = [LayerNorm(n_feat_out), ReLU(inplace=True)]
            if n_feat_proj_in != n_feat_out:
                mods.append(Linear(n_feat_out, n_feat_proj_in))
            query_proj = _NormReluProj(*mods)
        graph_conv = GATv2Conv(n_feat_proj_in, n_feat_agg // n_heads, heads=n_heads, add_self_loops=False)
        out_proj = Linear(n_feat_agg, n_feat_out) if n_feat_agg != n_feat_out else None
        norm_pre_mlp = LayerNorm(n_feat_out) if use_norm_pre_mlp else None
        mlp = get_linear_layers((2 + n_hidden) * [n_feat_out], init_activation=False, final_activation=False, norm=False)
        return n_feat_agg, query_proj, graph_conv, out_proj, norm_pre_mlp, mlp

    @staticmethod
    def _run(proj_features, graph_wrapper, prev, query_proj, graph_conv, out_proj, norm_pre_mlp, mlp,
             projected_sources=None):
        query = None if prev is None else query_proj(prev)
        plan = plan_for(graph_wrapper, proj_features)
        x = graph_conv.aggregate(proj_features.values, query, plan, projected_sources=projected_sources)
        if out_proj is not None:
            x = out_proj(x)
        if prev is not None:
            assert x.shape == prev.shape
            x = prev + x
        h = x
        if norm_pre_mlp is not None:
            h = norm_pre_mlp.ln_relu(h)
        return x + mlp(h)


class Proj2View(_AxialAttentionUpdate):
    """Resection layer: projection features -> view features."""

    def __init__(self, n_feat_proj_in, n_feat_view_out, n_heads, stateful=True, use_norm_pre_mlp=True,
                 n_feat_proj2view_agg=None, n_hidden_layers_view_update=0):
        super().__init__()
        self.n_feat_view_out = n_feat_view_out
        agg, q, conv, out_proj, norm, mlp = self._init_axial(
            n_feat_proj_in, n_feat_view_out, n_heads, stateful, use_norm_pre_mlp, n_feat_proj2view_agg,
            n_hidden_layers_view_update)
        self.n_feat_proj2view_agg = agg
        if q is not None:
            self.norm_and_proj_view2proj = q
        self.graph_conv = conv
        if out_proj is not None:
            self.proj_proj2view = out_proj
        if norm is not None:
            self.norm_pre_mlp = norm
        self.mlp = mlp

    def forward(self, proj_features, graph_wrapper, prev_view_features=None, projected_sources=None):
        assert self.stateful == (prev_view_features is not None)
        return self._run(proj_features, graph_wrapper, prev_view_features,
                         getattr(self, "norm_and_proj_view2proj", None), self.graph_conv,
                         getattr(self, "proj_proj2view", None), getattr(self, "norm_pre_mlp", None), self.mlp,
                         projected_sources)


class Proj2ScenePoint(_AxialAttentionUpdate):
    """Intersection layer: projection features -> scenepoint features."""

    def __init__(self, n_feat_proj_in, n_feat_scenepoint_out, n_heads, stateful=True, use_norm_pre_mlp=True,
                 n_feat_proj2scenepoint_agg=None, n_hidden_layers_scenepoint_update=0):
        super().__init__()
        self.n_feat_scenepoint_out = n_feat_scenepoint_out
        agg, q, conv, out_proj, norm, mlp = self._init_axial(
            n_feat_proj_in, n_feat_scenepoint_out, n_heads, stateful, use_norm_pre_mlp,
            n_feat_proj2scenepoint_agg, n_hidden_layers_scenepoint_update)
        self.n_feat_proj2scenepoint_agg = agg
        if q is not None:
            self.norm_and_proj_scenepoint2proj = q
        self.graph_conv = conv
        if out_proj is not None:
            self.proj_proj2scenepoint = out_proj
        if norm is not None:
            self.norm_pre_mlp = norm
        self.mlp = mlp

    def forward(self, proj_features, graph_wrapper, prev_scenepoint_features=None, projected_sources=None):
        assert self.stateful == (prev_scenepoint_features is not None)
        return self._run(proj_features, graph_wrapper, prev_scenepoint_features,
                         getattr(self, "norm_and_proj_scenepoint2proj", None), self.graph_conv,
                         getattr(self, "proj_proj2scenepoint", None), getattr(self, "norm_pre_mlp", None), self.mlp,
                         projected_sources)


class ViewAndScenePoint2Global(Module):
    """Global aggregation from view features and scenepoint features (layers.py:460-603): two
    single-target GATv2 graphs whose sources are the valid views / tracks."""

    def __init__(self, n_feat_scenepoint_in, n_feat_view_in, n_feat_global_out, n_heads, stateful=True,
                 use_norm_pre_mlp=True, n_feat_scenepoint2global_agg=None, n_feat_view2global_agg=None,
                 n_hidden_layers_global_update=0):
        super().__init__()
        self.n_feat_scenepoint_in = n_feat_scenepoint_in
        self.n_feat_view_in = n_feat_view_in
        self.n_feat_global_out = n_feat_global_out
        self.stateful = stateful
        self.use_norm_pre_mlp = use_norm_pre_mlp
        self.n_feat_scenepoint2global_agg = (_round_up(n_feat_scenepoint_in, n_heads)
                                             if n_feat_scenepoint2global_agg is None else n_feat_scenepoint2global_agg)
        self.n_feat_view2global_agg = (_round_up(n_feat_view_in, n_heads)
                                       if n_feat_view2global_agg is None else n_feat_view2global_agg)
        assert self.n_feat_scenepoint2global_agg % n_heads == 0
        assert self.n_feat_view2global_agg % n_heads == 0

        def query_proj(width):
            mods = [LayerNorm(n_feat_global_out), ReLU(inplace=True)]
            if width != n_feat_global_out:
                mods.append(Linear(n_feat_global_out, width))
            return _NormReluProj(*mods)

        if stateful:
            self.norm_and_proj_global2view = query_proj(n_feat_view_in)
        self.graph_conv_view2global = GATv2Conv(n_feat_view_in, self.n_feat_view2global_agg // n_heads,
                                                heads=n_heads, add_self_loops=False)
        if stateful:
            self.norm_and_proj_global2scenepoint = query_proj(n_feat_scenepoint_in)
        self.graph_conv_scenepoint2global = GATv2Conv(n_feat_scenepoint_in, self.n_feat_scenepoint2global_agg // n_heads,
                                                      heads=n_heads, add_self_loops=False)
        if (self.n_feat_view2global_agg + self.n_feat_scenepoint2global_agg) != n_feat_global_out:
            self.proj_view_and_scenepoint2global = Linear(
                self.n_feat_view2global_agg + self.n_feat_scenepoint2global_agg, n_feat_global_out)
        if use_norm_pre_mlp:
            self.norm_pre_mlp = LayerNorm(n_feat_global_out)
        self.mlp = get_linear_layers((2 + n_hidden_layers_global_update) * [n_feat_global_out],
                                     init_activation=False, final_activation=False, norm=False)

    def forward(self, view_features, scenepoint_features, graph_wrapper_view2global,
                graph_wrapper_scenepoint2global, prev_global_features=None):
        assert self.stateful == (prev_global_features is not None)
        qv = qp = None
        if prev_global_features is not None:
            qv = self.norm_and_proj_global2view(prev_global_features)
            qp = self.norm_and_proj_global2scenepoint(prev_global_features)
        v = self.graph_conv_view2global.aggregate(view_features, qv, plan_for(graph_wrapper_view2global))
        s = self.graph_conv_scenepoint2global.aggregate(scenepoint_features, qp, plan_for(graph_wrapper_scenepoint2global))
        x = torch.cat((v, s), dim=1)
        assert x.shape == (1, self.n_feat_view2global_agg + self.n_feat_scenepoint2global_agg)
        if hasattr(self, "proj_view_and_scenepoint2global"):
            x = self.proj_view_and_scenepoint2global(x)
        if prev_global_features is not None:
            assert x.shape == prev_global_features.shape
            x = prev_global_features + x
        h = x
        if self.use_norm_pre_mlp:
            h = self.norm_pre_mlp.ln_relu(h)
        return x + self.mlp(h)


class _Global2Node(Module):
    """x + [mlp](lin_node(ReLU(LN(x))) + lin_global(ReLU(LN(g)))) (layers.py:605-721)."""

    def _init_g2n(self, n_feat_global_in, width, n_hidden, use_norm):
        node_norm = LayerNorm(width) if use_norm else None
        global_norm = LayerNorm(n_feat_global_in) if use_norm else None
        lin_node = Linear(width, width)
        lin_global = Linear(n_feat_global_in, width, bias=False)
        mlp = None
        if n_hidden > 0:
            mlp = get_linear_layers(n_hidden * [width] + [width], init_activation=False, final_activation=False, norm=False)
        return node_norm, global_norm, lin_node, lin_global, mlp

    @staticmethod
    def _run(prev, glob, node_norm, global_norm, lin_node, lin_global, mlp):
        x = prev if node_norm is None else node_norm.ln_relu(prev)
        g = glob if global_norm is None else global_norm.ln_relu(glob)
        x = lin_node(x) + lin_global(g)
        if mlp is not None:
            x = mlp(F.relu(x))
        return prev + x


class Global2View(_Global2Node):
    def __init__(self, n_feat_global_in, n_feat_view_in_out, n_hidden_layers_view_update=0,
                 use_norm_global2view_update=True):
        super().__init__()
        self.n_feat_global_in, self.n_feat_view_in_out = n_feat_global_in, n_feat_view_in_out
        self.n_hidden_layers_view_update = n_hidden_layers_view_update
        self.use_norm_global2view_update = use_norm_global2view_update
        nn_, gn, ln, lg, mlp = self._init_g2n(n_feat_global_in, n_feat_view_in_out, n_hidden_layers_view_update,
                                              use_norm_global2view_update)
        if nn_ is not None:
            self.view_norm_layer, self.global_norm_layer = nn_, gn
        self.lin_view, self.lin_global = ln, lg
        if mlp is not None:
            self.mlp = mlp

    def forward(self, global_features, prev_view_features):
        return self._run(prev_view_features, global_features, getattr(self, "view_norm_layer", None),
                         getattr(self, "global_norm_layer", None), self.lin_view, self.lin_global,
                         getattr(self, "mlp", None))


class Global2ScenePoint(_Global2Node):
    def __init__(self, n_feat_global_in, n_feat_scenepoint_in_out, n_hidden_layers_scenepoint_update=0,
                 use_norm_global2scenepoint_update=True):
        super().__init__()
        self.n_feat_global_in, self.n_feat_scenepoint_in_out = n_feat_global_in, n_feat_scenepoint_in_out
        self.n_hidden_layers_scenepoint_update = n_hidden_layers_scenepoint_update
        self.use_norm_global2scenepoint_update = use_norm_global2scenepoint_update
        nn_, gn, ln, lg, mlp = self._init_g2n(n_feat_global_in, n_feat_scenepoint_in_out,
                                              n_hidden_layers_scenepoint_update, use_norm_global2scenepoint_update)
        if nn_ is not None:
            self.scenepoint_norm_layer, self.global_norm_layer = nn_, gn
        self.lin_scenepoint, self.lin_global = ln, lg
        if mlp is not None:
            self.mlp = mlp

    def forward(self, global_features, prev_scenepoint_features):
        return self._run(prev_scenepoint_features, global_features, getattr(self, "scenepoint_norm_layer", None),
                         getattr(self, "global_norm_layer", None), self.lin_scenepoint, self.lin_global,
                         getattr(self, "mlp", None))


class GraphAttnSfMGlobalFeatureUpdate(Module):
    def __init__(
        self,
        n_feat_proj_in,
        n_feat_scenepoint_out,
        n_feat_view_out,
        n_feat_proj2scenepoint_agg = None,
        n_feat_proj2view_agg = None,
        n_feat_global_out = None,
        n_feat_scenepoint2global_agg = None,
        n_feat_view2global_agg = None,
        output_global = True,
        n_heads = 1,
        stateful = True,
        global2view_and_global2scenepoint_enabled = True,
        n_hidden_layers_scenepoint_update = 0,
        n_hidden_layers_view_update = 0,
        n_hidden_layers_global_update = 0,
    ):
        super().__init__()
        self.n_feat_proj_in = n_feat_proj_in
        self.n_feat_scenepoint_out = n_feat_scenepoint_out
        self.n_feat_view_out = n_feat_view_out
        self.n_feat_global_out = n_feat_global_out
        self.global2view_and_global2scenepoint_enabled = global2view_and_global2scenepoint_enabled
        self.output_global = output_global
        needs_global = output_global or global2view_and_global2scenepoint_enabled
        if needs_global:
            assert n_feat_global_out is not None and n_feat_global_out % n_heads == 0
        assert n_feat_scenepoint_out % n_heads == 0
        assert n_feat_view_out % n_heads == 0
        self.proj2view = Proj2View(
            n_feat_proj_in, n_feat_view_out, n_heads, stateful=stateful, use_norm_pre_mlp=True,
            n_feat_proj2view_agg=n_feat_proj2view_agg, n_hidden_layers_view_update=n_hidden_layers_view_update)
        self.proj2scenepoint = Proj2ScenePoint(
            n_feat_proj_in, n_feat_scenepoint_out, n_heads, stateful=stateful, use_norm_pre_mlp=True,
            n_feat_proj2scenepoint_agg=n_feat_proj2scenepoint_agg,
            n_hidden_layers_scenepoint_update=n_hidden_layers_scenepoint_update)
        if needs_global:
            self.view_and_scenepoint2global = ViewAndScenePoint2Global(
                n_feat_scenepoint_out, n_feat_view_out, n_feat_global_out, n_heads, stateful=stateful,
                use_norm_pre_mlp=True, n_feat_scenepoint2global_agg=n_feat_scenepoint2global_agg,
                n_feat_view2global_agg=n_feat_view2global_agg,
                n_hidden_layers_global_update=n_hidden_layers_global_update)
        if global2view_and_global2scenepoint_enabled:
            self.global2view = Global2View(n_feat_global_out, n_feat_view_out,
                                           n_hidden_layers_view_update=n_hidden_layers_view_update,
                                           use_norm_global2view_update=True)
            self.global2scenepoint = Global2ScenePoint(n_feat_global_out, n_feat_scenepoint_out,
                                                       n_hidden_layers_scenepoint_update=n_hidden_layers_scenepoint_update,
                                                       use_norm_global2scenepoint_update=True)

    def forward(self, x, graph_structure, prev_scenepoint_features=None, prev_view_features=None,
                prev_global_features=None, projected_sources=None):
        """``projected_sources`` (extension): (lin_l(x) of proj2scenepoint, lin_l(x) of proj2view) when the caller
        already computed them; otherwise they are computed here (together, as one autograd node)."""
        m, n, n_feat_proj_in = x.shape
        assert n_feat_proj_in == self.n_feat_proj_in
        if projected_sources is None:
            conv_sp, conv_v = self.proj2scenepoint.graph_conv, self.proj2view.graph_conv
            projected_sources = ops.linear_multi(x.values, [(conv_sp.lin_l.weight, conv_sp.lin_l.bias),
                                                            (conv_v.lin_l.weight, conv_v.lin_l.bias)])
        scenepoint_features = self.proj2scenepoint(x, graph_structure['proj2scenepoint'],
                                                   prev_scenepoint_features=prev_scenepoint_features,
                                                   projected_sources=projected_sources[0])
        assert scenepoint_features.shape == (n, self.n_feat_scenepoint_out)
        view_features = self.proj2view(x, graph_structure['proj2view'], prev_view_features=prev_view_features,
                                       projected_sources=projected_sources[1])
        assert view_features.shape == (m, self.n_feat_view_out)
        global_features = None
        if self.output_global or self.global2view_and_global2scenepoint_enabled:
            global_features = self.view_and_scenepoint2global(
                view_features, scenepoint_features, graph_structure['view2global'],
                graph_structure['scenepoint2global'], prev_global_features=prev_global_features)
            assert global_features.shape == (1, self.n_feat_global_out)
        if self.global2view_and_global2scenepoint_enabled:
            scenepoint_features = self.global2scenepoint(global_features, scenepoint_features)
            view_features = self.global2view(global_features, view_features)
        if not self.output_global:
            return scenepoint_features, view_features
        return scenepoint_features, view_features, global_features


class GraphAttnSfMProjectionFeatureUpdate(Module):
    def __init__(self, n_feat_proj_in, n_feat_scenepoint_in, n_feat_view_in, n_feat_global_in, n_feat_proj_out,
                 n_hidden_layers_proj_update=0, normalize_global_features=True):
        super().__init__()
        self.n_feat_proj_in = n_feat_proj_in
        self.n_feat_scenepoint_in = n_feat_scenepoint_in
        self.n_feat_view_in = n_feat_view_in
        self.n_feat_global_in = n_feat_global_in
        self.n_feat_proj_out = n_feat_proj_out
        self.n_hidden_layers_proj_update = n_hidden_layers_proj_update
        self.normalize_global_features = normalize_global_features
        if normalize_global_features:
            self.scenepoint_norm_layer = LayerNorm(n_feat_scenepoint_in)
            self.view_norm_layer = LayerNorm(n_feat_view_in)
            self.global_norm_layer = LayerNorm(n_feat_global_in)
        self.lin_proj = Linear(n_feat_proj_in, n_feat_proj_out)
        self.lin_scenepoint = Linear(n_feat_scenepoint_in, n_feat_proj_out, bias=False)
        self.lin_view = Linear(n_feat_view_in, n_feat_proj_out, bias=False)
        self.lin_global = Linear(n_feat_global_in, n_feat_proj_out, bias=False)
        if n_hidden_layers_proj_update > 0:
            self.mlp = get_linear_layers(n_hidden_layers_proj_update * [n_feat_proj_out] + [n_feat_proj_out],
                                         init_activation=False, final_activation=False, norm=False)

    def forward(self, scenepoint_features, view_features, global_features, x, residual=None, projected=None, rowmax_slot=None):
        """(lin_proj(x) + lin_sp(sp)[col] + lin_view(view)[row] + lin_global(g)) / 4 [-> relu -> mlp]
        (layers.py:911-956).  Extensions: ``residual`` = SparseMat added to the result inside the same
        kernel (the skip connection of GraphAttnSfMLayer, layers.py:254-261); ``projected`` = the already
        computed 0.25 * lin_proj over the main feature block of ``x``."""
        if self.normalize_global_features:
            scenepoint_features = self.scenepoint_norm_layer.ln_relu(scenepoint_features)
            view_features = self.view_norm_layer.ln_relu(view_features)
            global_features = self.global_norm_layer.ln_relu(global_features)
        sp = self.lin_scenepoint(scenepoint_features)
        view = self.lin_view(view_features)
        glob = self.lin_global(global_features)

        w, b = self.lin_proj.weight, self.lin_proj.bias
        x0 = w0 = None
        if isinstance(x, _LazyFeatureCat) and x._extra.shape[2] <= 4:
            d_main = x._main.shape[2]
            proj = projected if projected is not None else ops.linear(x._main.values, 0.25 * w[:, :d_main], 0.25 * b)
            x0, w0 = x._extra.values, w[:, d_main:]
        elif isinstance(x, _LazyFeatureCat):
            proj = ops.linear(x.values, 0.25 * w, 0.25 * b)        # wide init features: materialised cat
        else:
            proj = projected if projected is not None else ops.linear(x.values, 0.25 * w, 0.25 * b)
        has_mlp = self.n_hidden_layers_proj_update > 0
        fused_skip = None if (residual is None or has_mlp) else residual.values
        new = ops.edge_update(proj, x0, w0, sp, view, glob, fused_skip, index_for(x), 1.0, 0.25,
                              rowmax_slot=rowmax_slot if projected is not None else None)
        assert new.shape == (x.indices.shape[1], self.n_feat_proj_out)
        if has_mlp:
            new = self.mlp(F.relu(new))
            if residual is not None:
                new = residual.values + new
        return x.with_values(new) if not isinstance(x, _LazyFeatureCat) else x._main.with_values(new)


class ProjLayer(Module):
    def __init__(self, n_feat_proj_in, n_feat_proj_out):
        super().__init__()
        self.lin_proj = Linear(n_feat_proj_in, n_feat_proj_out)

    def forward(self, x):
        return x.with_values(ops.linear(x.values, self.lin_proj.weight, self.lin_proj.bias))


def normalize_projection_features(x, norm_layer=None):
    """LayerNorm over the feature axis of the observation features, or mean-centering over the
    observations when no layer is given (layers.py:972-979)."""
    if norm_layer is not None:
        return x.with_values(F.layer_norm(x.values, norm_layer.normalized_shape, norm_layer.weight,
                                          norm_layer.bias, norm_layer.eps))
    return x.with_values(x.values - x.values.mean(dim=0, keepdim=True))


def relu_on_projection_features(x, _fused_norm=None):
    """ReLU on the observation features (layers.py:982-984).  ``_fused_norm=(x, LayerNorm)`` runs the
    preceding LayerNorm in the same kernel (normalize_projection_features + relu, layers.py:232-234)."""
    norm = None
    if _fused_norm is not None:
        x, norm = _fused_norm
    if not ops.ln_relu_width_supported(x.values.shape[1]):
        # widths the fused kernel does not take (> 256 and not a multiple of 4): plain torch
        v = x.values if norm is None else F.layer_norm(x.values, norm.normalized_shape, norm.weight, norm.bias, norm.eps)
        return x.with_values(F.relu(v))
    if norm is None:
        return x.with_values(ops.ln_relu(x.values))
    return x.with_values(ops.ln_relu(x.values, norm.weight, norm.bias, norm.eps))


class IdentityLayer(Module):
    def forward(self, x):
        return x


class EmbeddingLayer(Module):
    def __init__(self, pos_emb_n_freq, in_dim, post_embed_proj_dim=None):
        super().__init__()
        if pos_emb_n_freq > 0:
            self.embed, self.d_out = get_embedder(pos_emb_n_freq, in_dim)
        else:
            self.embed, self.d_out = (Identity(), in_dim)
        self.post_embed_proj_dim = post_embed_proj_dim
        if post_embed_proj_dim is not None:
            if post_embed_proj_dim == -1:
                post_embed_proj_dim = self.d_out
            self.post_embed_lin = Linear(self.d_out, post_embed_proj_dim)
            self.d_out = post_embed_proj_dim
        else:
            self.post_embed_lin = None

    def forward(self, x):
        feats = self.embed(x.values)
        if self.post_embed_lin is not None:
            feats = self.post_embed_lin(feats)
        return x.with_values(feats)
