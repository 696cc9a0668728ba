"""``GATv2Conv`` with the parameter layout of ``torch_geometric.nn.GATv2Conv`` (``att [1,H,C]``,
``bias [H*C]``, ``lin_l``, ``lin_r``) and the fused sm_100a edge kernel as its message passing.

The reference constructs it as ``GATv2Conv(in, C, heads=H, add_self_loops=False)``
(``code/models/layers.py:304-309,401-406,506-511,521-526``).  Only the element rows are ever
sources and only the aggregation rows are ever targets (``code/utils/dataset_utils.py:511-576``), so
``lin_l`` runs on the E element rows alone and ``lin_r`` on the T aggregation rows alone; the
skipped halves receive zero gradient in the reference too, so results and gradients are identical.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from .. import ops
from ..index import plan_from_targets


class GATv2Conv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, edge_dim=None, fill_value="mean", bias=True, share_weights=False):
        super().__init__()
        if add_self_loops or edge_dim is not None or share_weights or not concat or dropout != 0.0 \
                or negative_slope != ops.LEAKY_SLOPE or not bias:
            raise NotImplementedError("gasfm_b200.GATv2Conv implements the configuration GASFM uses: "
                                      "add_self_loops=False, concat=True, slope 0.2, no dropout / edge features")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope = negative_slope
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, heads * out_channels, bias=True)
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        """PyG initialisation: glorot on the last two dims for weights and ``att``, uniform
        +-1/sqrt(in) for the linear biases, zeros for the output bias."""
        with torch.no_grad():
            for w in (self.lin_l.weight, self.lin_r.weight, self.att):
                a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
                w.uniform_(-a, a)
            bound = 1.0 / math.sqrt(self.in_channels)
            self.lin_l.bias.uniform_(-bound, bound)
            self.lin_r.bias.uniform_(-bound, bound)
            self.bias.zero_()

    def project_sources(self, x_elements):
        """lin_l on the source rows: tcgen05 3xTF32 GEMM for observation-sized inputs (ops.linear)."""
        return ops.linear(x_elements, self.lin_l.weight, self.lin_l.bias)

    def project_targets(self, x_agg):
        """Query projection; ``None`` = zero query features -> one broadcast row ``lin_r.bias``."""
        if x_agg is None:
            # zero query features: x_r = 0 @ W^T + b.  The (exactly zero) weight term keeps lin_r.weight in
            # the autograd graph so that it receives the same all-zero gradient it gets in the reference
            # (whose training loop concatenates every parameter's .grad, code/train.py:137).
            return (self.lin_r.bias + self.lin_r.weight.sum() * 0.0).unsqueeze(0)
        return ops.linear(x_agg, self.lin_r.weight, self.lin_r.bias)

    def aggregate(self, x_elements, x_agg, plan, projected_sources=None):
        """[T, H*C] attention-aggregate of the element rows over ``plan``'s segments.  ``projected_sources``:
        lin_l(x_elements) when the caller already computed it, optionally as ``(tensor, lazy, slot)`` where ``lazy()``
        rebuilds the tensor in backward instead of keeping it (activation recompute) and ``slot`` is where the backward
        kernel leaves the row maxima of dXL for the block's input-gradient GEMM (ops.EdgeBlockContext)."""
        lazy = slot = None
        if isinstance(projected_sources, tuple):
            projected_sources, lazy, slot = (tuple(projected_sources) + (None, None))[:3]
        xl = self.project_sources(x_elements) if projected_sources is None else projected_sources
        shard = getattr(plan, "shard", None)
        if shard is not None and shard.world > 1:
            # track-sharded scene: this rank holds only part of every segment -> merge across ranks
            from .. import dist as gdist
            return gdist.sharded_gat(xl, self.project_targets(x_agg), self.att, self.bias, plan, self.heads,
                                     shard.exchange, lazy_xl=lazy, rowmax_slot=slot)
        return ops.gat_edge_attention(xl, self.project_targets(x_agg), self.att, self.bias, plan, self.heads, lazy_xl=lazy,
                                      rowmax_slot=slot)

    def forward(self, x, edge_index):
        """PyG call convention ``conv(x[N,d], edge_index[2,E]) -> [N, H*C]`` for arbitrary graphs
        (API parity; GASFM's own layers call ``aggregate`` with the scene's cached plans)."""
        n_nodes = x.shape[0]
        src, dst = edge_index[0], edge_index[1]
        plan = plan_from_targets(dst, n_nodes)
        xl = self.project_sources(x).index_select(0, src)
        return ops.gat_edge_attention(xl, self.project_targets(x), self.att, self.bias, plan, self.heads)
