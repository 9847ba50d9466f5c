"""Localise a product-vs-oracle gradient mismatch: per rank, the gradient of H_U before the gradient exchange
(`grad_u<layer>`, inner / halo part) and of H after it (`grad_h<layer>`), from ``Buffer.trace`` on the CUDA side and
``OracleRank.trace`` on the oracle side.  The default arguments reproduce the ReLU-kink case documented in
tests/test_parity_gpu.py::test_training_parity_eight_partitions (graph_seed 0: rank 0, row 64).

    python tests/localize_gradient_mismatch.py [--parts 8] [--graph-seed 0] [--rate 0.5] [--hidden 32] [--epochs 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import torch  # noqa: E402
from tests.harness import make_args, _relerr  # noqa: E402
from bns_gcn_b200.data import make_graph, partition_graph  # noqa: E402
from bns_gcn_b200 import train  # noqa: E402
from bns_gcn_b200.helper import context as ctx  # noqa: E402
from bns_gcn_b200.helper.comm import run_threads  # noqa: E402
from oracle import bns_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--parts", type=int, default=8)
ap.add_argument("--graph-seed", type=int, default=0)
ap.add_argument("--rate", type=float, default=0.5)
ap.add_argument("--hidden", type=int, default=32)
ap.add_argument("--epochs", type=int, default=2)
a = ap.parse_args()
P, E = a.parts, a.epochs
fg = make_graph("small", seed=a.graph_seed)
parts = partition_graph(fg, P, "random", seed=a.graph_seed)
args = make_args(dataset="small", model="graphsage", sampling_rate=a.rate, n_layers=3, n_hidden=a.hidden, n_partitions=P)


def prod_fn(comm, r):
    p = parts[r]
    ar = argparse.Namespace(**vars(args))
    ar.n_feat, ar.n_class, ar.n_train = p.meta["n_feat"], p.meta["n_class"], p.meta["n_train"]
    st = train.setup(p.graph, p.node_dict, p.gpb, ar, "cuda:0")
    buf, tr, sel = ctx.buffer._get(), {}, []
    for e in range(E):
        buf.trace = tr if e == E - 1 else None
        train.train_epoch(st, e)
        sel.append([None if s is None else s.cpu().clone() for s in st.selected])
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in tr.items()}, sel


prod = run_threads(P, prod_fn, device="cuda:0")
sel = [[prod[r][1][e] for r in range(P)] for e in range(E)]


def orc_fn(comm, r):
    rk = O.OracleRank(O.RankInput.from_partition(parts[r]), comm, model="graphsage", n_layers=3, n_hidden=a.hidden,
                      sampling_rate=a.rate, dropout=0.0, seed=0)
    for e in range(E):
        rk.epoch(selected=sel[e][r], trace=True)
    return {k: v for k, v in rk.trace.items() if k.startswith("grad_")}


orc = O.run_threads(P, orc_fn)
for r in range(P):
    tp, to = prod[r][0], orc[r]
    n_in = to["grad_h2"].shape[0]
    line = []
    for k in sorted(to.keys()):
        x, y = tp[k], to[k]
        if k.startswith("grad_u"):
            line.append(f"{k}[inner] {_relerr(x[:n_in], y[:n_in]):.1e} [halo] {_relerr(x[n_in:], y[n_in:]):.1e}")
        else:
            line.append(f"{k} {_relerr(x, y):.1e}")
    print("rank", r, " | ".join(line), flush=True)
    for k in sorted(to.keys()):
        d = (tp[k] - to[k]).norm(dim=1) / to[k].norm().clamp(min=1e-30)
        if float(d.max()) > 1e-4:
            top = torch.topk(d, 4)
            print(f"    {k}: worst rows {top.indices.tolist()} {[float(f'{v:.1e}') for v in top.values.tolist()]}")
