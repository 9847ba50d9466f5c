"""Dense layers of the path (K8, module/layer.py:30, 38, 83, 92 of the reference are plain ``nn.Linear`` in fp32).

The reference runs them as true-fp32 cuBLAS SGEMMs (torch 1.12: ``allow_tf32=False`` for matmul).  On B200 the fp32
SIMT pipe gives ~45 TFLOP/s -- after the SpMM work it is the largest share of the epoch -- while one TF32 tensor-core
pass would miss the 1e-4 parity bar (10-bit mantissa).  ``linear()`` below therefore uses the error-compensated
**3xTF32** scheme: split every f32 operand into ``hi = tf32(x)`` and ``lo = x - hi`` (exact in f32) and accumulate
``hi*hi + hi*lo + lo*hi`` in three tensor-core GEMMs with f32 accumulation; the dropped ``lo*lo`` term is 2^-22
relative, i.e. f32-level accuracy.  The GEMMs themselves are library calls (cuBLAS TF32, what the contract allows for
plain GEMMs); a hand-written tcgen05 kernel with the LayerNorm / ReLU / dropout epilogue fused is the SURVEY §8(f)
rank-2 follow-up.

``MODE`` (env BNS_DENSE): "fp32" (default, the reference precision) | "3xtf32".  Measured on B200 (round 1, Reddit
shape, N=1): 3xtf32 built from three cuBLAS TF32 GEMMs + the split passes is SLOWER than fp32 SIMT cuBLAS (44.2 vs
37.8 ms/epoch), so it is off by default; the win needs the split fused into a hand-written tcgen05 kernel.
"""
import os

import torch
import torch.nn.functional as F

MODE = os.environ.get("BNS_DENSE", "fp32")


def _split(t: torch.Tensor):
    """hi = t rounded to TF32 (10 explicit mantissa bits, round-to-nearest on the 13 dropped bits), lo = t - hi."""
    bits = t.contiguous().view(torch.int32)
    hi = ((bits + 0x1000) & -0x2000).view(torch.float32)
    return hi, t - hi


def _mm3(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b with 3xTF32 error compensation (a: [m, k], b: [k, n], both f32)."""
    ah, al = _split(a)
    bh, bl = _split(b)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        out = torch.mm(al, bh)          # small terms first, then the dominant one: better rounding
        out.addmm_(ah, bl)
        out.addmm_(ah, bh)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


class _Linear3x(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        y = _mm3(x, weight.t())
        if bias is not None:
            y += bias
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = _mm3(dy, weight) if ctx.needs_input_grad[0] else None
        dw = _mm3(dy.t(), x) if ctx.needs_input_grad[1] else None
        db = dy.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """Drop-in for ``F.linear`` on 2-D f32 CUDA inputs."""
    if MODE == "3xtf32" and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2:
        return _Linear3x.apply(x, weight, bias)
    return F.linear(x, weight, bias)
