"""Pure-torch restatement of ``torch_geometric.nn.GATv2Conv`` (oracle, test-only).

The reference builds its attention layers from PyG's ``GATv2Conv`` with
``heads=H, add_self_loops=False`` and otherwise default arguments
(``/root/reference/code/models/layers.py:304-309,401-406,506-511,521-526``) and calls
them as ``conv(x, edge_index)`` on a node matrix of ``E`` element rows followed
by ``T`` aggregation rows (``layers.py:329-335,426-432,550-556,566-572``).

PyG is not vendored in the reference and not installable offline, so the
published algorithm (Brody et al., "How attentive are graph attention
networks?"; PyG 2.2 ``GATv2Conv`` with ``concat=True, negative_slope=0.2,
dropout=0, bias=True, share_weights=False, edge_dim=None``) is restated here:

    x_l = lin_l(x).view(N, H, C)            (source transform)
    x_r = lin_r(x).view(N, H, C)            (target transform)
    e_ji = sum_c att[h, c] * leaky_relu(x_l[j] + x_r[i], 0.2)[h, c]
    alpha_ji = softmax over {j : j -> i} of e_ji   (max-subtracted, denom + 1e-16)
    out[i] = concat_h sum_j alpha_ji * x_l[j]  + bias

The module keeps PyG's parameter names and shapes (``att [1,H,C]``, ``bias
[H*C]``, ``lin_l.{weight,bias}``, ``lin_r.{weight,bias}``) so state_dicts are
interchangeable with the reference's.  It deliberately materialises every
``[E,H,C]`` intermediate with index_select / scatter, like PyG does, so that it
is also a fair stand-in for the reference's CPU cost.
"""
import math

import torch
from torch import nn


def glorot_(t):
    """PyG ``glorot``: U(-a, a), a = sqrt(6 / (size(-2) + size(-1)))."""
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


def gatv2_edge_softmax_aggregate(x_l, x_r, att, edge_index, negative_slope=0.2):
    """The message-passing core.  x_l, x_r: [N,H,C]; att: [1,H,C] -> [N,H,C]."""
    src, dst = edge_index[0], edge_index[1]
    n_nodes, n_heads, _ = x_l.shape
    x_j = x_l.index_select(0, src)
    x_i = x_r.index_select(0, dst)
    e = torch.nn.functional.leaky_relu(x_i + x_j, negative_slope)
    score = (e * att).sum(dim=-1)  # [E,H]
    # PyG softmax(): subtract the per-target max, exponentiate, divide by sum + 1e-16.
    seg_max = score.new_full((n_nodes, n_heads), float("-inf"))
    seg_max = seg_max.scatter_reduce(
        0, dst[:, None].expand(-1, n_heads), score.detach(), "amax", include_self=True
    )
    p = (score - seg_max.index_select(0, dst)).exp()
    denom = score.new_zeros((n_nodes, n_heads)).index_add(0, dst, p)
    alpha = p / (denom.index_select(0, dst) + 1e-16)
    msg = x_j * alpha.unsqueeze(-1)
    return torch.zeros_like(x_l).index_add(0, dst, msg)


class GATv2Conv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0.0, add_self_loops=True,
                 edge_dim=None, fill_value="mean", bias=True, share_weights=False):
        super().__init__()
        if add_self_loops or edge_dim is not None or share_weights or not concat or dropout != 0.0:
            raise NotImplementedError("only the configuration the reference uses is restated")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.heads = heads
        self.negative_slope = negative_slope
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=bias)
        self.lin_r = nn.Linear(in_channels, heads * out_channels, bias=bias)
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.lin_l.weight)
        glorot_(self.lin_r.weight)
        glorot_(self.att)
        with torch.no_grad():
            if self.lin_l.bias is not None:
                # PyG ``Linear`` default bias init: U(-1/sqrt(in), 1/sqrt(in)).
                bound = 1.0 / math.sqrt(self.in_channels)
                self.lin_l.bias.uniform_(-bound, bound)
                self.lin_r.bias.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.zero_()

    def forward(self, x, edge_index):
        h, c = self.heads, self.out_channels
        x_l = self.lin_l(x).view(-1, h, c)
        x_r = self.lin_r(x).view(-1, h, c)
        out = gatv2_edge_softmax_aggregate(x_l, x_r, self.att, edge_index, self.negative_slope)
        out = out.reshape(-1, h * c)
        if self.bias is not None:
            out = out + self.bias
        return out
