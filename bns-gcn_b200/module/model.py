"""``GCN`` / ``GraphSAGE`` layer stacks with the reference's constructor and ``forward`` signatures
(module/model.py:7-93).  ``GAT`` and ``--norm batch`` are SURVEY.md §8(f) "next" rows."""
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..helper import context as ctx
from .layer import GCNLayer, GraphSAGELayer

# LayerNorm -> ReLU -> (next layer's) dropout as one fused kernel each way (ops.LnReluDropout) when the model uses
# `--norm layer` with ReLU on CUDA; False = the three separate ATen ops of the reference (module/model.py:88-91, :80)
FUSE_NORM_ACT_DROPOUT = True


class GNNBase(nn.Module):

    def __init__(self, layer_size, activation, use_pp=False, dropout=0.5, norm='layer', n_linear=0):
        super(GNNBase, self).__init__()
        self.n_layers = len(layer_size) - 1
        self.layers = nn.ModuleList()
        self.activation = activation
        self.use_pp = use_pp
        self.n_linear = n_linear
        if norm is None:
            self.use_norm = False
        else:
            self.use_norm = True
            self.norm = nn.ModuleList()
        self.dropout = nn.Dropout(p=dropout)

    def _build(self, layer_cls, layer_size, use_pp, norm, train_size):
        for i in range(self.n_layers):
            if i < self.n_layers - self.n_linear:
                self.layers.append(layer_cls(layer_size[i], layer_size[i + 1], use_pp=use_pp))
            else:
                self.layers.append(nn.Linear(layer_size[i], layer_size[i + 1]))
            if i < self.n_layers - 1 and self.use_norm:
                if norm == 'layer':
                    self.norm.append(nn.LayerNorm(layer_size[i + 1], elementwise_affine=True))
                elif norm == 'batch':
                    from .sync_bn import SyncBatchNorm
                    self.norm.append(SyncBatchNorm(layer_size[i + 1], train_size))
            use_pp = False                                   # model.py:40, 75: only layer 0 is precomputed

    def _forward(self, g, feat, *norms):
        h = feat
        dropped = False                      # this layer's input dropout was already applied by the fused op
        for i in range(self.n_layers):
            if not dropped:
                h = self.dropout(h)
            dropped = False
            if i < self.n_layers - self.n_linear:
                if self.training and (i > 0 or not self.use_pp):
                    h = ctx.buffer.update(i, h, overlap=True)          # model.py:47-48, 82-83
                h = self.layers[i](g, h, *norms)
            else:
                h = self.layers[i](h)
            if i < self.n_layers - 1:
                nm = self.norm[i] if self.use_norm else None
                if (FUSE_NORM_ACT_DROPOUT and isinstance(nm, nn.LayerNorm) and self.activation is F.relu
                        and nm.elementwise_affine and ops.ln_relu_dropout_supported(h, h.shape[1])):
                    p = self.dropout.p if self.training else 0.0
                    h = ops.LnReluDropout.apply(h, nm.weight, nm.bias, nm.eps, p, ops.RNG["seed"] + 7919 * (i + 1))
                    dropped = True
                else:
                    if self.use_norm:
                        h = nm(h)
                    h = self.activation(h)
        return h


class GCN(GNNBase):

    def __init__(self, layer_size, activation, use_pp, dropout=0.5, norm='layer', train_size=None, n_linear=0):
        super(GCN, self).__init__(layer_size, activation, use_pp, dropout, norm, n_linear)
        self._build(GCNLayer, layer_size, use_pp, norm, train_size)

    def forward(self, g, feat, in_norm=None, out_norm=None):
        return self._forward(g, feat, in_norm, out_norm)


class GraphSAGE(GNNBase):

    def __init__(self, layer_size, activation, use_pp, dropout=0.5, norm='layer', train_size=None, n_linear=0):
        super(GraphSAGE, self).__init__(layer_size, activation, use_pp, dropout, norm, n_linear)
        self._build(GraphSAGELayer, layer_size, use_pp, norm, train_size)

    def forward(self, g, feat, in_norm=None):
        return self._forward(g, feat, in_norm)


class GAT(GNNBase):
    """module/model.py:96-132."""

    def __init__(self, layer_size, activation, use_pp, heads=1, dropout=0.5, norm='layer', train_size=None, n_linear=0):
        super(GAT, self).__init__(layer_size, activation, use_pp, dropout, norm, n_linear)
        from .gat import GATConv
        for i in range(self.n_layers):
            if i < self.n_layers - self.n_linear:
                self.layers.append(GATConv(layer_size[i], layer_size[i + 1], heads, dropout, dropout))
            else:
                self.layers.append(nn.Linear(layer_size[i], layer_size[i + 1]))
            if i < self.n_layers - 1 and self.use_norm:
                if norm == 'layer':
                    self.norm.append(nn.LayerNorm(layer_size[i + 1], elementwise_affine=True))
                elif norm == 'batch':
                    from .sync_bn import SyncBatchNorm
                    self.norm.append(SyncBatchNorm(layer_size[i + 1], train_size))

    def forward(self, g, feat):
        h = feat
        for i in range(self.n_layers):
            if i < self.n_layers - self.n_linear:
                if self.training:
                    if i > 0 or not self.use_pp:
                        h1 = ctx.buffer.update(i, h, overlap=True)                # model.py:117-118
                    else:
                        h1 = h
                        h = h[0:g.num_nodes('_V')]                                # :120-121
                    h = self.layers[i](g, (h1, h))
                else:
                    h = self.layers[i](g, h)
                h = h.mean(1)
            else:
                h = self.dropout(h)
                h = self.layers[i](h)
            if i < self.n_layers - 1:
                if self.use_norm:
                    h = self.norm[i](h)
                h = self.activation(h)
        return h
