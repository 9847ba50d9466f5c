"""SpMM microbenchmark: bns_spmm_sum_f32 vs torch.sparse.mm (cuSPARSE CSR SpMM, what DGL 0.9's
gspmm('copy_lhs','sum') dispatches to) on a BASELINE-shaped graph.  CUDA-event timing, L2 flushed between
iterations.  Prints one JSON line per case.

  python tools/bench_spmm.py --shape reddit --parts 1 --F 256 [--chunk 1024] [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bns_gcn_b200  # noqa: E402
from bns_gcn_b200 import ops  # noqa: E402
from bns_gcn_b200.data import make_graph, partition_graph  # noqa: E402

HBM_PEAK = 6576.7e9


def time_it(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--parts", type=int, default=1)
    ap.add_argument("--F", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-cusparse", action="store_true")
    ap.add_argument("--transpose", action="store_true")
    ap.add_argument("--slab", type=int, default=0)
    ap.add_argument("--col-blocks", type=int, default=1,
                    help="experiment: split the source rows into this many blocks, one accumulate pass per block")
    ap.add_argument("--n", type=int, default=0, help="override node count")
    ap.add_argument("--e", type=int, default=0, help="override edge count")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    ov = {}
    if a.n:
        ov["n"] = a.n
    if a.e:
        ov["e"] = a.e
    fg = make_graph(a.shape, with_feat=False, **ov)
    if a.parts == 1:
        indptr, idx, n_src = fg.indptr, fg.src, fg.n_nodes
    else:
        p = partition_graph(fg, a.parts, "random", ranks=[0])[0]
        indptr, idx, n_src = p.graph.indptr, p.graph.indices, p.graph.num_nodes()
    n_dst, nnz = indptr.numel() - 1, idx.numel()
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), n_src, a.chunk)
    if a.transpose:
        g = g.transpose()
        n_dst, n_src = g.n_rows, g.n_cols
    x = torch.randn(n_src, a.F, device=dev)
    y = torch.empty(n_dst, a.F, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if a.col_blocks > 1:
        ip, ix = g.csr()
        rows = torch.repeat_interleave(torch.arange(g.n_rows, device=dev), ip[1:] - ip[:-1])
        blocks, B = [], a.col_blocks
        for b in range(B):
            c0, c1 = (n_src * b) // B, (n_src * (b + 1)) // B
            m = (ix >= c0) & (ix < c1)
            cnt = torch.bincount(rows[m], minlength=g.n_rows)
            ipb = torch.zeros(g.n_rows + 1, dtype=torch.int64, device=dev)
            ipb[1:] = torch.cumsum(cnt, 0)
            blocks.append((ops.DeviceGraph.from_csr(ipb, (ix[m] - c0).int(), c1 - c0, a.chunk), c0, c1))
        del rows

        def run_blocked():
            for i, (gb, c0, c1) in enumerate(blocks):
                ops.spmm(gb, x[c0:c1], y, accumulate=i > 0, slab=a.slab)
        y_ref = ops.spmm(g, x, slab=a.slab).clone()
        run_blocked()
        print("blocked vs plain rel err", ((y - y_ref).norm() / y_ref.norm()).item(), file=sys.stderr)
        med, best = time_it(run_blocked, a.iters, flush)
    else:
        med, best = time_it(lambda: ops.spmm(g, x, y, slab=a.slab), a.iters, flush)
    bytes_alg = 8 * (n_dst + 1) + 4 * nnz + 4 * a.F * n_src + 4 * a.F * n_dst
    bytes_gather = 8 * (n_dst + 1) + 4 * nnz + 4 * a.F * nnz + 4 * a.F * n_dst
    res = {"case": f"{a.shape}/P{a.parts}/F{a.F}/slab{a.slab}/chunk{a.chunk}/colblocks{a.col_blocks}" + ("/T" if a.transpose else ""), "n_dst": n_dst, "n_src": n_src,
           "nnz": nnz, "chunks": g.n_chunks, "split_rows": g.n_split_rows, "ms_median": round(med, 4),
           "ms_best": round(best, 4), "alg_GBs": round(bytes_alg / med / 1e6, 1),
           "alg_frac_of_hbm": round(bytes_alg / (med * 1e-3) / HBM_PEAK, 4),
           "gather_GBs": round(bytes_gather / med / 1e6, 1)}
    if not a.no_cusparse:
        csr = torch.sparse_csr_tensor(indptr.to(dev), idx.to(dev), torch.ones(nnz, device=dev), size=(n_dst, n_src)) \
            if not a.transpose else None
        if csr is not None:
            ref = torch.sparse.mm(csr, x)
            err = ((ref - y).norm() / ref.norm()).item()
            cmed, cbest = time_it(lambda: torch.sparse.mm(csr, x), max(3, a.iters // 2), flush)
            res.update({"cusparse_ms_median": round(cmed, 4), "speedup_vs_cusparse": round(cmed / med, 2),
                        "relerr_vs_cusparse": err})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
