"""``GraphAttnSfMNet``: mirror of the reference's ``code/models/graph_attn_sfm.py`` (constructor
reads the same ``model.*`` keys, :12-41; ``forward(data) -> {"Ps_norm", "pts3D"[, "depths"]}``, :117-185;
identical module tree / state_dict names) running on the sm_100a kernels."""
import torch
from torch import nn

from .baseNet import BaseNet
from .layers import (EmbeddingLayer, GraphAttnSfMGlobalFeatureUpdate, GraphAttnSfMLayer, get_linear_layers,
                     relu_on_projection_features)
from .. import ops
from ..index import _INDEX_ATTR, ObservationIndex
from ..utils.sparse_utils import SparseMat


class GraphAttnSfMNet(BaseNet):
    def __init__(self, conf, batchnorm=False):
        super().__init__(conf)
        g = lambda key, **kw: conf.get_int('model.' + key, **kw)  # noqa: E731
        b = lambda key, **kw: conf.get_bool('model.' + key, **kw)  # noqa: E731
        num_layers = g('num_layers')
        n_heads = g('n_heads')
        n_feat_proj = self.n_feat_proj = g('n_feat_proj')
        n_feat_scenepoint = g('n_feat_scenepoint')
        n_feat_view = g('n_feat_view')
        n_feat_global = g('n_feat_global')
        agg = dict(n_feat_proj2scenepoint_agg=g('n_feat_proj2scenepoint_agg', default=None),
                   n_feat_proj2view_agg=g('n_feat_proj2view_agg', default=None),
                   n_feat_scenepoint2global_agg=g('n_feat_scenepoint2global_agg', default=None),
                   n_feat_view2global_agg=g('n_feat_view2global_agg', default=None))
        hidden = dict(n_hidden_layers_scenepoint_update=g('n_hidden_layers_scenepoint_update'),
                      n_hidden_layers_view_update=g('n_hidden_layers_view_update'),
                      n_hidden_layers_global_update=g('n_hidden_layers_global_update'))
        n_hidden_layers_proj_update = g('n_hidden_layers_proj_update')
        pos_emb_n_freq = g('pos_emb_n_freq')
        self.use_norm_proj_update = b('use_norm_proj_update')
        add_residual_skipconn_proj_update = b('add_residual_skipconn_proj_update')
        self.add_skipconn_from_init_projfeat = b('add_skipconn_from_init_projfeat')
        self.stateful_global_features = b('stateful_global_features')
        g2n = b('global2view_and_global2scenepoint_enabled')
        self.depth_head_enabled = b('depth_head.enabled', default=False)
        self.view_head_enabled = b('view_head.enabled', default=False)
        self.scenepoint_head_enabled = b('scenepoint_head.enabled', default=False)
        self.batchnorm = batchnorm
        if batchnorm:
            raise NotImplementedError()

        self.embed = EmbeddingLayer(pos_emb_n_freq, 2, post_embed_proj_dim=-1)
        d_emb = self.embed.d_out
        self.n_feat_skipconn_init_projfeat_in = d_emb if self.add_skipconn_from_init_projfeat else 0
        last_width = g('depth_head.n_feat') if self.depth_head_enabled else n_feat_proj

        self.equivariant_blocks = nn.ModuleList()
        for i in range(num_layers):
            first = i == 0
            self.equivariant_blocks.append(GraphAttnSfMLayer(
                d_emb if first else n_feat_proj,
                last_width if i == num_layers - 1 else n_feat_proj,
                n_feat_scenepoint, n_feat_view, n_feat_global,
                use_norm_proj_update=self.use_norm_proj_update,
                add_residual_skipconn_proj_update=add_residual_skipconn_proj_update,
                n_feat_skipconn_init_projfeat_in=(self.n_feat_skipconn_init_projfeat_in
                                                  if (not first and self.add_skipconn_from_init_projfeat) else None),
                n_heads=n_heads,
                stateful=False if first else self.stateful_global_features,
                global2view_and_global2scenepoint_enabled=g2n,
                n_hidden_layers_proj_update=n_hidden_layers_proj_update,
                **agg, **hidden))

        if self.view_head_enabled or self.scenepoint_head_enabled:
            if not self.view_head_enabled and self.scenepoint_head_enabled:
                raise NotImplementedError('Final feature aggregation for only view features or scenepoint '
                                          'features alone is not implemented.')
            self.final_global_update = GraphAttnSfMGlobalFeatureUpdate(
                last_width, n_feat_scenepoint, n_feat_view, n_feat_global_out=n_feat_global, output_global=False,
                n_heads=n_heads, stateful=self.stateful_global_features,
                global2view_and_global2scenepoint_enabled=g2n, **agg, **hidden)
        if self.depth_head_enabled:
            self.depth_head = get_linear_layers((1 + g('depth_head.n_hidden_layers')) * [last_width] + [1],
                                                init_activation=False, final_activation=False, norm=False)
        if self.view_head_enabled:
            self.view_head = get_linear_layers((1 + g('view_head.n_hidden_layers')) * [n_feat_view] + [self.out_channels],
                                               init_activation=False, final_activation=False, norm=False)
        if self.scenepoint_head_enabled:
            self.scenepoint_head = get_linear_layers((1 + g('scenepoint_head.n_hidden_layers')) * [n_feat_scenepoint] + [3],
                                                     init_activation=False, final_activation=False, norm=False)

    @staticmethod
    def _observations(data):
        """``data.x`` as a gasfm_b200 SparseMat sharing the scene's cached CSR/CSC index.  Accepts the
        reference's own ``SparseMat`` (duck-typed: values / indices / cam_per_pts / pts_per_cam / shape)."""
        x = data.x
        if not x.values.is_cuda:
            raise RuntimeError("gasfm_b200.GraphAttnSfMNet runs on CUDA tensors only: call data.to(device) first "
                               "(there is no CPU fallback)")
        idx = getattr(x, _INDEX_ATTR, None) or getattr(data, _INDEX_ATTR, None)
        if idx is None or idx.device != x.indices.device or idx.n_obs != x.indices.shape[1]:
            idx = ObservationIndex(x.indices, x.shape[0], x.shape[1])
            for holder in (x, data):
                try:
                    setattr(holder, _INDEX_ATTR, idx)
                except AttributeError:
                    pass
        idx.shard = getattr(data, "shard", None) or getattr(x, "shard", None)     # track-sharded scene (gasfm_b200.dist)
        return SparseMat(x.values, x.indices, x.cam_per_pts, x.pts_per_cam, tuple(x.shape), _index=idx)

    def _recompute_plan(self, n_obs, device):
        """Number of leading blocks that recompute their activations in backward (policy: ``ops.recompute_plan``)."""
        return ops.recompute_plan(n_obs, self.n_feat_proj, len(self.equivariant_blocks),
                                  torch.cuda.get_device_properties(device).total_memory, torch.is_grad_enabled())

    def forward(self, data):
        graph_structure = data.graph_wrappers
        observations = self._observations(data)
        n_recompute = self._recompute_plan(observations.indices.shape[1], observations.values.device)
        ops.last_recompute_plan = (n_recompute, len(self.equivariant_blocks))
        ops.set_activation_recompute(n_recompute > 0, first_of_forward=True)
        projection_features = self.embed(observations)   # [m,n,2] -> [m,n,d_emb]
        if not self.use_norm_proj_update:
            # in-place ReLU aliasing of the reference (layers.py:982-984): without a norm layer, block 0
            # rectifies the embedding tensor that later blocks concatenate as the init skip connection
            projection_features = relu_on_projection_features(projection_features)
        skipconn = projection_features if self.add_skipconn_from_init_projfeat else None
        scenepoint_features = view_features = global_features = None
        stateful = self.stateful_global_features
        for i_block, block in enumerate(self.equivariant_blocks):
            ops.set_activation_recompute(i_block < n_recompute)
            projection_features, scenepoint_features, view_features, global_features = block(
                projection_features, graph_structure,
                prev_scenepoint_features=scenepoint_features if stateful else None,
                prev_view_features=view_features if stateful else None,
                prev_global_features=global_features if stateful else None,
                skipconn_init_projfeat=skipconn)

        pred_dict = {}
        if self.depth_head_enabled:
            n_views, n_scenepoints = projection_features.shape[:2]
            depth = projection_features.with_values(self.depth_head(projection_features.values))
            pred_dict.update(self.extract_depth_outputs(depth))
        if self.view_head_enabled or self.scenepoint_head_enabled:
            n_input, m_input = self.final_global_update(
                projection_features, graph_structure,
                prev_scenepoint_features=scenepoint_features if stateful else None,
                prev_view_features=view_features if stateful else None,
                prev_global_features=global_features if stateful else None)
            if self.view_head_enabled:
                pred_dict.update(self.extract_view_outputs(self.view_head(torch.relu(m_input))))
            if self.scenepoint_head_enabled:
                pred_dict.update(self.extract_scenepoint_outputs(self.scenepoint_head(torch.relu(n_input)).T))
        return pred_dict
