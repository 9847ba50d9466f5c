"""``SyncBatchNorm`` (``--norm batch``): batch normalisation whose statistics span ALL partitions -- the layer behind
``module/sync_bn.py`` of the reference, with its class surface (``SyncBatchNorm(num_features, whole_size, eps,
momentum)``, parameters ``weight`` / ``bias``, buffers ``running_mean`` / ``running_var``: state dicts interchange) and
its conventions: column sums run over every inner node of every rank and are divided by ``whole_size`` (the global
number of TRAIN nodes, module/model.py:39 of the reference), the variance is the one-pass ``(S2 - mean S1) / n``, and
the weight / bias gradients come out as global sums (the Reducer divides them by n_train and all-reduces them once more,
like any other parameter).

How it runs is different.  Per layer and step the reference launches ~12 element-wise / reduction ATen kernels and four
``[F]`` all-reduces (sync_bn.py:17-18, :35-36).  Here (csrc/fused.cuh):

    forward    ``bns_bn_colsums_f32(mode 0)``  one pass -> ``[sum x | sum x^2]``           (deterministic two-stage sum)
               one packed ``[2F]`` all-reduce
               ``bns_bn_apply_f32``            one pass: statistics -> normalise -> affine, running stats, saves mean/rstd
    backward   ``bns_bn_colsums_f32(mode 1)``  one pass -> ``[sum dy | sum dy x_hat]``     (x_hat recomputed, not stored)
               one packed ``[2F]`` all-reduce  -> these are d(bias), d(weight)
               ``bns_bn_bwd_f32``              one pass -> dx
"""
import torch
from torch import nn

from .._lib import check, lib
from ..helper import context as ctx

_WS = {}


def _scratch(F: int, device) -> torch.Tensor:
    key = (F, device, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = torch.empty(lib.bns_bn_workspace_bytes(F), dtype=torch.uint8, device=device)
    return ws


def _kernels_take(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 4 == 0
            and x.shape[1] <= 1024 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0)


def _colsums(mode: int, a: torch.Tensor, x, mean, rstd) -> torch.Tensor:
    n, F = a.shape
    out = torch.empty(2 * F, dtype=torch.float32, device=a.device)
    ws = _scratch(F, a.device)
    with torch.cuda.device(a.device):
        check(lib.bns_bn_colsums_f32(mode, a.data_ptr(), a.stride(0), None if x is None else x.data_ptr(),
                                     0 if x is None else x.stride(0), n, F, None if mean is None else mean.data_ptr(),
                                     None if rstd is None else rstd.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                     torch.cuda.current_stream(a.device).cuda_stream), "bns_bn_colsums_f32")
    return out


class _SyncBN(torch.autograd.Function):
    """Training-mode forward / backward on the CUDA kernels; ``comm.all_reduce_sum`` carries the packed moments."""

    @staticmethod
    def forward(ctx_, x, weight, bias, module: "SyncBatchNorm", comm):
        x = x.contiguous()
        n, F = x.shape
        sums = _colsums(0, x, None, None, None)
        comm.all_reduce_sum(sums)
        y = torch.empty_like(x)
        mean = torch.empty(F, dtype=torch.float32, device=x.device)
        rstd = torch.empty(F, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.bns_bn_apply_f32(x.data_ptr(), x.stride(0), n, F, sums.data_ptr(), float(module.whole_size),
                                       float(module.eps), weight.data_ptr(), bias.data_ptr(), float(module.momentum),
                                       module.running_mean.data_ptr(), module.running_var.data_ptr(), y.data_ptr(),
                                       y.stride(0), mean.data_ptr(), rstd.data_ptr(),
                                       torch.cuda.current_stream(x.device).cuda_stream), "bns_bn_apply_f32")
        ctx_.save_for_backward(x, weight, mean, rstd)
        ctx_.n, ctx_.comm = float(module.whole_size), comm
        return y

    @staticmethod
    def backward(ctx_, dy):
        x, weight, mean, rstd = ctx_.saved_tensors
        dy = dy.contiguous()
        n, F = x.shape
        sums = _colsums(1, dy, x, mean, rstd)                 # [d bias | d weight] of this rank
        ctx_.comm.all_reduce_sum(sums)
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(lib.bns_bn_bwd_f32(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), n, F, mean.data_ptr(),
                                     rstd.data_ptr(), weight.data_ptr(), sums.data_ptr(), ctx_.n, dx.data_ptr(),
                                     dx.stride(0), torch.cuda.current_stream(x.device).cuda_stream), "bns_bn_bwd_f32")
        return dx, sums[F:], sums[:F], None, None


class SyncBatchNorm(nn.Module):

    def __init__(self, num_features, whole_size, eps=1e-5, momentum=0.1):
        super().__init__()
        self.whole_size, self.eps, self.momentum = whole_size, eps, momentum
        self.register_buffer('running_mean', torch.zeros(num_features))
        self.register_buffer('running_var', torch.ones(num_features))
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))

    def forward(self, x):
        if not self.training:
            # inference: a per-column affine map built from the running statistics
            scale = self.weight * torch.rsqrt(self.running_var + self.eps)
            return torch.addcmul(self.bias - self.running_mean * scale, x, scale)
        if not _kernels_take(x):
            raise NotImplementedError("SyncBatchNorm: training needs a row-major f32 CUDA matrix whose width is a multiple "
                                      "of 4 and at most 1024 (there is no CPU path)")
        return _SyncBN.apply(x, self.weight, self.bias, self, ctx.comm())
