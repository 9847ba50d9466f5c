"""``SyncBatchNorm`` (``--norm batch``; reference: module/sync_bn.py:7-56): batch statistics over ALL partitions.

Same arithmetic as the reference, including its conventions: sums run over every inner node of every rank but are
divided by ``whole_size`` (= the global number of TRAIN nodes, model.py:39), and the returned dweight / dbias are
already global sums (the Reducer then divides by n_train and all-reduces them once more, like any other parameter).
The reference issues 4 all-reduces of ``[F]`` per layer (2 forward, 2 backward, :17-18, :35-36); here each pair
travels as one ``[2F]`` message."""
import torch
from torch import nn
from torch.autograd import Function

from ..helper import context as ctx


class SyncBatchNormFunc(Function):

    @staticmethod
    def forward(ctx_, x, weight, bias, whole_size, running_mean, running_var, training, momentum, eps, comm):
        if not training:
            mean, var = running_mean, running_var
        else:
            F = x.shape[1]
            s = torch.cat([x.sum(dim=0), (x ** 2).sum(dim=0)])
            comm.all_reduce_sum(s)                                   # sync_bn.py:17-18 in one message
            sum_x, sum_x2 = s[:F], s[F:]
            mean = sum_x / whole_size
            var = (sum_x2 - mean * sum_x) / whole_size
            running_mean.mul_(1 - momentum).add_(mean * momentum)
            running_var.mul_(1 - momentum).add_(var * momentum)
        std = torch.sqrt(var + eps)
        x_hat = (x - mean) / std
        if training:
            ctx_.save_for_backward(x_hat, weight, std)
            ctx_.whole_size, ctx_.comm = whole_size, comm
        return x_hat * weight + bias

    @staticmethod
    def backward(ctx_, grad):
        x_hat, weight, std = ctx_.saved_tensors
        F = grad.shape[1]
        d = torch.cat([grad.sum(dim=0), (grad * x_hat).sum(dim=0)])
        ctx_.comm.all_reduce_sum(d)                                  # sync_bn.py:35-36 in one message
        dbias, dweight = d[:F], d[F:]
        n = ctx_.whole_size
        dx = (weight / n) / std * (n * grad - dbias - x_hat * dweight)
        return dx, dweight, dbias, None, None, None, None, None, None, None


class SyncBatchNorm(nn.Module):

    def __init__(self, num_features, whole_size, eps=1e-5, momentum=0.1):
        super(SyncBatchNorm, self).__init__()
        self.register_buffer('running_mean', torch.zeros(num_features))
        self.register_buffer('running_var', torch.ones(num_features))
        self.whole_size = whole_size
        self.eps = eps
        self.momentum = momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))

    def forward(self, x):
        return SyncBatchNormFunc.apply(x, self.weight, self.bias, self.whole_size, self.running_mean,
                                       self.running_var, self.training, self.momentum, self.eps, ctx.comm())
