"""``SetOfSetNet`` (DPESFM baseline): mirror of the reference's ``code/models/SetOfSet.py`` on the
row / column mean-pooling kernels.  Same module tree and state_dict names."""
import torch
from torch import nn

from .baseNet import BaseNet
from .layers import (EmbeddingLayer, ProjLayer, SetOfSetGlobalFeatureUpdate, SetOfSetLayer, get_linear_layers,
                     normalize_projection_features, relu_on_projection_features)
from .graph_attn_sfm import GraphAttnSfMNet


class SetOfSetBlock(nn.Module):
    def __init__(self, d_in, d_out, conf):
        super().__init__()
        self.block_size = conf.get_int("model.block_size")
        self.proj_feat_normalization = conf.get_bool("model.proj_feat_normalization")
        self.add_skipconn_for_residual_blocks = conf.get_bool("model.add_skipconn_for_residual_blocks")
        self.layers = nn.ModuleList([SetOfSetLayer(d_in if i == 0 else d_out, d_out) for i in range(self.block_size)])
        if self.add_skipconn_for_residual_blocks:
            self.skip_projection = None if d_in == d_out else ProjLayer(d_in, d_out)

    def forward(self, x):
        xl = x
        for i, layer in enumerate(self.layers):
            xl = layer(xl)
            if i < len(self.layers) - 1:
                if self.proj_feat_normalization:
                    xl = normalize_projection_features(xl)
                xl = relu_on_projection_features(xl)
        if self.add_skipconn_for_residual_blocks:
            x_skip = x
            if self.skip_projection is not None:
                x_skip = self.skip_projection(x_skip)
                if self.proj_feat_normalization:
                    x_skip = normalize_projection_features(x_skip)
            xl = x_skip + xl
        return relu_on_projection_features(xl)


class SetOfSetNet(BaseNet):
    def __init__(self, conf, batchnorm=False):
        super().__init__(conf)
        num_blocks = conf.get_int('model.num_blocks')
        num_feats = conf.get_int('model.num_features')
        pos_emb_n_freq = conf.get_int('model.pos_emb_n_freq')
        self.depth_head_enabled = conf.get_bool('model.depth_head.enabled', default=False)
        self.view_head_enabled = conf.get_bool('model.view_head.enabled', default=False)
        self.scenepoint_head_enabled = conf.get_bool('model.scenepoint_head.enabled', default=False)
        if batchnorm:
            raise NotImplementedError()
        last_width = conf.get_int("model.depth_head.n_feat") if self.depth_head_enabled else num_feats
        self.embed = EmbeddingLayer(pos_emb_n_freq, 2)
        self.equivariant_blocks = nn.ModuleList([
            SetOfSetBlock(self.embed.d_out if i == 0 else num_feats, last_width if i == num_blocks - 1 else num_feats, conf)
            for i in range(num_blocks)])
        if self.view_head_enabled or self.scenepoint_head_enabled:
            if not self.view_head_enabled and self.scenepoint_head_enabled:
                raise NotImplementedError()
            self.final_global_update = SetOfSetGlobalFeatureUpdate(num_feats, num_feats, output_global=False)
        if self.depth_head_enabled:
            self.depth_head = get_linear_layers((1 + conf.get_int('model.depth_head.n_hidden_layers')) * [last_width] + [1],
                                                init_activation=False, final_activation=False, norm=False)
        if self.view_head_enabled:
            self.view_head = get_linear_layers((1 + conf.get_int('model.view_head.n_hidden_layers')) * [num_feats] + [self.out_channels],
                                               init_activation=False, final_activation=False, norm=False)
        if self.scenepoint_head_enabled:
            self.scenepoint_head = get_linear_layers((1 + conf.get_int('model.scenepoint_head.n_hidden_layers')) * [num_feats] + [3],
                                                     init_activation=False, final_activation=False, norm=False)

    def forward(self, data):
        x = self.embed(GraphAttnSfMNet._observations(data))
        for block in self.equivariant_blocks:
            x = block(x)
        pred_dict = {}
        if self.depth_head_enabled:
            pred_dict.update(self.extract_depth_outputs(x.with_values(self.depth_head(x.values))))
        if self.view_head_enabled or self.scenepoint_head_enabled:
            n_input, m_input = self.final_global_update(x)
            if self.view_head_enabled:
                pred_dict.update(self.extract_view_outputs(self.view_head(torch.relu(m_input))))
            if self.scenepoint_head_enabled:
                pred_dict.update(self.extract_scenepoint_outputs(self.scenepoint_head(torch.relu(n_input)).T))
        return pred_dict
