"""``GCNLayer`` / ``GraphSAGELayer`` with the reference's constructor, parameter names, initialisation and
``forward`` signatures (module/layer.py:8-103); the DGL message passing inside is one call into libbnsgcn.so."""
import math

import torch
from torch import nn

from ..graph import FullGraphHandle, PartitionAggregate, PartitionGraph
from ..ops import AggregateSum


def _aggregate(graph, feat, rs, cs_u=None):
    """``update_all(copy_u, sum)`` with the row / column scalings fused (K1+K2)."""
    if isinstance(graph, PartitionGraph):
        cs_in = cs_halo = None
        if cs_u is not None:
            cs_in, cs_halo = cs_u[:graph.n_in], cs_u[graph.n_in:]
        return PartitionAggregate.apply(feat, graph, rs, cs_in, cs_halo, getattr(feat, '_bns_ready', None))
    if isinstance(graph, FullGraphHandle):
        return AggregateSum.apply(feat, graph.a, rs, cs_u)
    raise TypeError(f"unsupported graph handle {type(graph).__name__}")


class GCNLayer(nn.Module):

    def __init__(self, in_feats, out_feats, bias=True, use_pp=False):
        super(GCNLayer, self).__init__()
        self.use_pp = use_pp
        self.linear = nn.Linear(in_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.linear.weight.size(1))
        self.linear.weight.data.uniform_(-stdv, stdv)
        if self.linear.bias is not None:
            self.linear.bias.data.uniform_(-stdv, stdv)

    def forward(self, graph, feat, in_norm, out_norm):
        """``out_norm``: sqrt(out_deg) of every *local* node (inner then halo, static) -- the reference rebuilds a
        U-ordered copy of it every epoch (train.py:245-253), which the slot map makes unnecessary."""
        if self.training:
            if self.use_pp:
                return self.linear(feat)                                            # layer.py:29-30
            h = _aggregate(graph, feat, graph.recip(in_norm), graph.recip(out_norm))  # :32-38
            return self.linear(h)
        in_n = torch.sqrt(graph.in_degrees().float())                                # :40-45
        out_n = torch.sqrt(graph.out_degrees().float())
        return self.linear(_aggregate(graph, feat, 1.0 / in_n, 1.0 / out_n))


class GraphSAGELayer(nn.Module):

    def __init__(self, in_feats, out_feats, bias=True, use_pp=False):
        super(GraphSAGELayer, self).__init__()
        self.use_pp = use_pp
        if self.use_pp:
            self.linear = nn.Linear(2 * in_feats, out_feats, bias=bias)
        else:
            self.linear1 = nn.Linear(in_feats, out_feats, bias=bias)
            self.linear2 = nn.Linear(in_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        if self.use_pp:
            stdv = 1. / math.sqrt(self.linear.weight.size(1))
            self.linear.weight.data.uniform_(-stdv, stdv)
            if self.linear.bias is not None:
                self.linear.bias.data.uniform_(-stdv, stdv)
        else:
            stdv = 1. / math.sqrt(self.linear1.weight.size(1))
            self.linear1.weight.data.uniform_(-stdv, stdv)
            self.linear2.weight.data.uniform_(-stdv, stdv)
            if self.linear1.bias is not None:
                self.linear1.bias.data.uniform_(-stdv, stdv)
                self.linear2.bias.data.uniform_(-stdv, stdv)

    def forward(self, graph, feat, in_norm):
        if self.training:
            if self.use_pp:
                return self.linear(feat)                                            # layer.py:82-83
            num_dst = graph.num_nodes('_V')
            ah = _aggregate(graph, feat, graph.recip(in_norm))                       # :85-91  (sum / degs)
            return self.linear1(feat[0:num_dst]) + self.linear2(ah)                  # :92
        degs = graph.in_degrees()                                                    # :94-102
        ah = _aggregate(graph, feat, 1.0 / degs.float())
        if self.use_pp:
            return self.linear(torch.cat((feat, ah), dim=1))
        return self.linear1(feat) + self.linear2(ah)
