"""Where does the time of the split-GEMM dense modes go?  (diagnosis; prints ms per call)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bns_gcn_b200  # noqa
from bns_gcn_b200.module import dense

dev = torch.device("cuda:0")
M, K, N = 232965, 1204, 256


def t(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


x = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev)
dy = torch.randn(M, N, device=dev)
xb, wb, dyb = x.bfloat16(), w.bfloat16(), dy.bfloat16()
gf = 2 * M * K * N / 1e9
print(f"shape M={M} K={K} N={N}: {gf:.0f} GFLOP per GEMM")
r = {}
r["fp32 mm  x@w^T"] = t(lambda: torch.mm(x, w.t()))
r["fp32 mm  dy^T@x (dW)"] = t(lambda: torch.mm(dy.t(), x))
torch.backends.cuda.matmul.allow_tf32 = True
r["tf32 mm  x@w^T"] = t(lambda: torch.mm(x, w.t()))
r["tf32 mm  dy^T@x"] = t(lambda: torch.mm(dy.t(), x))
torch.backends.cuda.matmul.allow_tf32 = False
r["bf16 mm  x@w^T -> bf16"] = t(lambda: torch.mm(xb, wb.t()))
r["bf16 mm  x@w^T -> f32 (out_dtype)"] = t(lambda: torch.mm(xb, wb.t(), out_dtype=torch.float32))
acc = torch.mm(xb, wb.t(), out_dtype=torch.float32)
r["bf16 addmm(acc, x, w^T) -> f32"] = t(lambda: torch.addmm(acc, xb, wb.t(), out_dtype=torch.float32))
r["bf16 mm  dy^T@x -> f32"] = t(lambda: torch.mm(dyb.t(), xb, out_dtype=torch.float32))
r["split3 (x)"] = t(lambda: dense._split3(x))
r["tf32 split (x)"] = t(lambda: dense._split(x))
for k, v in r.items():
    print(f"{k:40s} {v:8.3f} ms   {gf / v:8.1f} TFLOP/s-equivalent")
