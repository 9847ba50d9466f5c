"""``SyncBatchNorm`` (``--norm batch``; reference: module/sync_bn.py:7-56): batch statistics over ALL partitions.

The arithmetic and its conventions are the reference's: sums run over every inner node of every rank but are divided
by ``whole_size`` (= the global number of TRAIN nodes, model.py:39); the returned d(weight) / d(bias) are already
global sums (the Reducer then divides by n_train and all-reduces them once more, like any other parameter).  What
differs is the traffic: the reference issues four ``[F]`` all-reduces per layer and step (sync_bn.py:17-18, :35-36);
here the two moments travel as one ``[2F]`` message forward and the two gradient sums as one backward."""
import torch
from torch import nn
from torch.autograd import Function

from ..helper import context as ctx


def _global_column_sums(comm, first: torch.Tensor, second: torch.Tensor):
    """Column sums of two ``[n, F]`` matrices over all ranks, one all-reduce for both."""
    width = first.shape[1]
    packed = torch.cat([first.sum(dim=0), second.sum(dim=0)])
    comm.all_reduce_sum(packed)
    return packed[:width], packed[width:]


class SyncBatchNormFunc(Function):
    """y = (x - mean) / sqrt(var + eps) * weight + bias with ``mean = S1 / n``, ``var = (S2 - mean * S1) / n``,
    ``S1 = sum x``, ``S2 = sum x^2`` over every rank, ``n = whole_size``."""

    @staticmethod
    def forward(ctx_, x, weight, bias, whole_size, running_mean, running_var, training, momentum, eps, comm):
        if training:
            s1, s2 = _global_column_sums(comm, x, x ** 2)
            mean = s1 / whole_size
            var = (s2 - mean * s1) / whole_size
            for running, batch in ((running_mean, mean), (running_var, var)):
                running.mul_(1 - momentum).add_(batch * momentum)
        else:
            mean, var = running_mean, running_var
        std = torch.sqrt(var + eps)
        x_hat = (x - mean) / std
        if training:
            ctx_.save_for_backward(x_hat, weight, std)
            ctx_.whole_size, ctx_.comm = whole_size, comm
        return x_hat * weight + bias

    @staticmethod
    def backward(ctx_, grad):
        x_hat, weight, std = ctx_.saved_tensors
        d_bias, d_weight = _global_column_sums(ctx_.comm, grad, grad * x_hat)
        n = ctx_.whole_size
        d_x = (weight / n) / std * (n * grad - d_bias - x_hat * d_weight)
        return (d_x, d_weight, d_bias) + (None,) * 7


class SyncBatchNorm(nn.Module):
    """Parameters ``weight`` / ``bias`` and buffers ``running_mean`` / ``running_var`` as in the reference, so that
    state dicts are interchangeable."""

    def __init__(self, num_features, whole_size, eps=1e-5, momentum=0.1):
        super().__init__()
        self.whole_size, self.eps, self.momentum = whole_size, eps, momentum
        for name, init in (('running_mean', torch.zeros), ('running_var', torch.ones)):
            self.register_buffer(name, init(num_features))
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))

    def forward(self, x):
        return SyncBatchNormFunc.apply(x, self.weight, self.bias, self.whole_size, self.running_mean,
                                       self.running_var, self.training, self.momentum, self.eps, ctx.comm())
