"""Generate golden vectors by running the UNMODIFIED reference (/root/reference: train.run, module/*, helper/*)
on CPU with the gloo backend, 2 processes, under a minimal stand-in for the `dgl` package.

Why a stand-in: the reference imports `dgl` / `ogb` at module top (helper/utils.py:4-8, module/layer.py:5) and
hard-codes `.cuda()` / `pin_memory=True`; neither wheel is installable here (no network) and there is no GPU in the
build container.  The shim below implements ONLY the DGL calls the path makes (heterograph, node_subgraph,
out_edges / out_degrees / remove_edges, update_all(copy_u, sum|mean), local_scope) with plain torch index ops, and
maps every `cuda` placement to the CPU.  Everything else -- get_boundary, get_pos, Buffer (gloo ring), Reducer,
precompute, construct_graph, the layers, the model, the epoch loop, Adam -- is the reference's own code, imported
from /root/reference and executed as is.  The only behavioural patch: the Reducer's thread pool runs its jobs at
`synchronize()` time instead of concurrently (on the GPU the jobs are slow enough that `param.grad` exists when they
touch it; on the CPU they would race with autograd).

    python tests/golden/make_golden.py          # writes tests/golden/ref_<model>_p2.pt

The inputs are regenerated from seeds by bns-gcn_b200/data (not stored); the file holds the reference's outputs.
This script is the only thing that reads /root/reference; tests only read the .pt files.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CONFIGS = {
    "graphsage": dict(shape="tiny", n_parts=2, model="graphsage", n_layers=3, n_hidden=16, rate=0.5, epochs=3),
    "gcn": dict(shape="tiny", n_parts=2, model="gcn", n_layers=3, n_hidden=16, rate=0.5, epochs=3),
    # the same two runs, plus the reference's evaluation forward on the whole graph with the trained weights
    "graphsage_eval": dict(shape="tiny", n_parts=2, model="graphsage", n_layers=3, n_hidden=16, rate=0.5, epochs=2,
                           eval_logits=True, slim=True),
    "gcn_eval": dict(shape="tiny", n_parts=2, model="gcn", n_layers=3, n_hidden=16, rate=0.5, epochs=2,
                     eval_logits=True, slim=True),
    "graphsage_bn": dict(shape="tiny", n_parts=2, model="graphsage", n_layers=3, n_hidden=16, rate=0.5, epochs=3,
                         norm="batch", graph_override={"train": 1.0}),   # whole_size == #nodes, as under --inductive
    # 8 gloo processes, 7 peers per rank; "slim": keep every rank's index sets but only rank 0's tensors (the reduced
    # gradients / weights are identical on all ranks) so that the fixture stays below 1 MB
    "graphsage_small": dict(shape="small", n_parts=8, model="graphsage", n_layers=3, n_hidden=32, rate=0.5, epochs=2,
                            slim=True),
    # GCN on the 6000-node graph over 4 ranks at a low sampling rate; GraphSAGE with a trailing nn.Linear layer
    # (--n-linear 1) on the inductive (train-nodes-only) graph over 3 ranks
    "gcn_small": dict(shape="small", n_parts=4, model="gcn", n_layers=3, n_hidden=32, rate=0.1, epochs=2, slim=True),
    "graphsage_nlin_induc": dict(shape="tiny", n_parts=3, model="graphsage", n_layers=3, n_hidden=16, rate=0.5, epochs=3,
                                 n_linear=1, inductive=True, slim=True),
    # GAT, 2 heads: the reference's GAT model / precompute / construct_feat / epoch loop (module/model.py:96-132,
    # train.py:208-209, :284-297, :401-402) run as they are; the layer they instantiate, dgl.nn.GATConv, comes from the
    # stand-in below (GATConvStandIn: DGL 0.9's published forward as a DENSE masked softmax -- a different formulation
    # from the oracle's and the product's entry-list one, so this pins the wiring AND cross-checks the op)
    "gat": dict(shape="tiny", n_parts=2, model="gat", n_layers=3, n_hidden=16, rate=0.5, epochs=3, heads=2),
    # --sampling-rate 0 (BNS-GCN's p = 0 end of the sweep): no boundary node is ever sampled, send sizes and ratios are 0
    # BASELINE configs[3] in miniature: multi-label targets, --dataset yelp selects BCEWithLogitsLoss(reduction='sum')
    # (train.py:358-361), 2-layer GAT, 1 head
    "gat_yelp": dict(shape="tiny-ml", n_parts=2, model="gat", n_layers=2, n_hidden=16, rate=0.5, epochs=3, heads=1,
                     dataset="yelp"),
    "graphsage_rate0": dict(shape="tiny", n_parts=2, model="graphsage", n_layers=3, n_hidden=16, rate=0.0, epochs=2,
                            slim=True),
}


# ------------------------------------------------------------------------------------------------------
# the dgl stand-in
# ------------------------------------------------------------------------------------------------------
class _Data(dict):
    pass


class _NodeView:
    def __init__(self):
        self.data = _Data()


class FakeGraph:
    """Homogeneous graph: the `subg` of load_partition and the in/out graphs derived from it."""

    def __init__(self, u, v, n):
        self.u, self.v, self.n = u.long(), v.long(), int(n)
        self.ndata, self.edata = _Data(), _Data()

    def num_nodes(self):
        return self.n

    def num_edges(self):
        return int(self.u.numel())

    def edges(self):
        return self.u, self.v

    def clone(self):
        g = FakeGraph(self.u.clone(), self.v.clone(), self.n)
        g.ndata.update(self.ndata)
        return g

    def int(self):
        return self

    def to(self, *_a, **_k):
        return self

    def out_edges(self, nodes, form="uv"):
        nodes = nodes.long()
        eids = []
        order = torch.argsort(self.u, stable=True)
        su = self.u[order]
        lo = torch.searchsorted(su, nodes)
        hi = torch.searchsorted(su, nodes, right=True)
        for a, b in zip(lo.tolist(), hi.tolist()):       # grouped in the order of `nodes`, by edge id inside a node
            eids.append(order[a:b])
        eid = torch.cat(eids) if eids else torch.empty(0, dtype=torch.long)
        if form == "eid":
            return eid
        return self.u[eid], self.v[eid]

    def out_degrees(self, nodes=None):
        deg = torch.bincount(self.u, minlength=self.n)
        return deg if nodes is None else deg[nodes.long()]

    def in_degrees(self):
        return torch.bincount(self.v, minlength=self.n)

    def remove_edges(self, eids):
        keep = torch.ones(self.u.numel(), dtype=torch.bool)
        keep[eids] = False
        self.u, self.v = self.u[keep], self.v[keep]

    # the evaluation branch of the layers (module/layer.py:39-45, 93-102) works on the homogeneous graph directly
    @contextlib.contextmanager
    def local_scope(self):
        saved = dict(self.ndata)
        try:
            yield
        finally:
            self.ndata.clear()
            self.ndata.update(saved)

    def update_all(self, msg, red):
        h = self.ndata[msg[1]]
        out = torch.zeros(self.n, *h.shape[1:], dtype=h.dtype).index_add_(0, self.v, h[self.u])
        if red[0] == "mean":
            out = out / self.in_degrees().clamp(min=1).to(h.dtype).view(-1, *([1] * (h.dim() - 1)))
        self.ndata[red[2]] = out


class _EdgeType:
    def __init__(self, g):
        self.g = g

    def update_all(self, msg, red, etype=None):
        g = self.g
        h = g.nodes["_U"].data[msg[1]]
        out = torch.zeros(g.n_v, *h.shape[1:], dtype=h.dtype).index_add_(0, g.v, h[g.u])
        if red[0] == "mean":
            cnt = torch.bincount(g.v, minlength=g.n_v).clamp(min=1).to(h.dtype)
            out = out / cnt.view(-1, *([1] * (h.dim() - 1)))
        g.nodes["_V"].data[red[2]] = out


class FakeHetero:
    """dgl.heterograph({('_U','_E','_V'): (u, v)}): node counts inferred from the largest id (train.py:276-279)."""

    def __init__(self, u, v):
        self.u, self.v = u.long(), v.long()
        self.n_u = int(self.u.max()) + 1 if self.u.numel() else 0
        self.n_v = int(self.v.max()) + 1 if self.v.numel() else 0
        self.nodes = {"_U": _NodeView(), "_V": _NodeView()}

    def num_nodes(self, ntype):
        return self.n_u if ntype == "_U" else self.n_v

    def add_nodes(self, n, ntype):
        if ntype == "_U":
            self.n_u += n
        else:
            self.n_v += n

    def __getitem__(self, key):
        return _EdgeType(self)

    @contextlib.contextmanager
    def local_scope(self):
        saved = {k: dict(v.data) for k, v in self.nodes.items()}
        try:
            yield
        finally:
            for k, v in self.nodes.items():
                v.data.clear()
                v.data.update(saved[k])


class GATConvStandIn(torch.nn.Module):
    """dgl.nn.GATConv (DGL 0.9, python/dgl/nn/pytorch/conv/gatconv.py) for the call the reference makes --
    ``GATConv(in, out, heads, feat_drop, attn_drop)`` applied as ``layer(g, (h_src, h_dst))`` on the bipartite graph:

        ft = fc(feat_drop(h)).view(-1, H, F);  el = (ft_src * attn_l).sum(-1);  er = (ft_dst * attn_r).sum(-1)
        e_uv = leaky_relu(el_u + er_v, 0.2);   a = attn_drop(edge_softmax(g, e));   rst_v = sum_u a_uv ft_u + bias

    Same constructor order / parameter names / initialisation as DGL (fc, attn_l, attn_r, bias; xavier-normal with the
    ReLU gain, zero bias).  The attention is written DENSE here: an [n_v, n_u, H] score tensor, -inf where there is no
    edge, softmax over the source axis."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2, residual=False,
                 activation=None, allow_zero_in_degree=False, bias=True):
        super().__init__()
        assert not residual and activation is None
        nn = torch.nn
        self._num_heads, self._out_feats = num_heads, out_feats
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.attn_r = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.feat_drop, self.attn_drop = nn.Dropout(feat_drop), nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.bias = nn.Parameter(torch.FloatTensor(size=(num_heads * out_feats,)))
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        nn.init.constant_(self.bias, 0)

    def forward(self, graph, feat):
        H, F = self._num_heads, self._out_feats
        h_src, h_dst = self.feat_drop(feat[0]), self.feat_drop(feat[1])
        ft_src = self.fc(h_src).view(-1, H, F)
        ft_dst = self.fc(h_dst).view(-1, H, F)
        n_u, n_v = ft_src.shape[0], ft_dst.shape[0]
        assert int(graph.u.max()) < n_u and int(graph.v.max()) < n_v
        if (torch.bincount(graph.v, minlength=n_v) == 0).any():
            raise RuntimeError("There are 0-in-degree nodes in the graph")       # DGLError in DGL
        el = (ft_src * self.attn_l).sum(dim=-1)                                  # [n_u, H]
        er = (ft_dst * self.attn_r).sum(dim=-1)                                  # [n_v, H]
        score = self.leaky_relu(el.unsqueeze(0) + er.unsqueeze(1))               # [n_v, n_u, H]
        adj = torch.zeros(n_v, n_u, dtype=torch.bool)
        adj[graph.v, graph.u] = True
        score = score.masked_fill(~adj.unsqueeze(-1), float('-inf'))
        a = self.attn_drop(torch.softmax(score, dim=1))
        rst = torch.einsum('vuh,uhf->vhf', a, ft_src)
        return rst + self.bias.view(1, H, F)


def _node_subgraph(g, mask):
    keep = torch.nonzero(mask, as_tuple=True)[0]
    new = torch.full((g.n,), -1, dtype=torch.long)
    new[keep] = torch.arange(keep.numel())
    ok = (new[g.u] >= 0) & (new[g.v] >= 0)
    return FakeGraph(new[g.u[ok]], new[g.v[ok]], keep.numel())


def install_dgl_shim():
    dgl = types.ModuleType("dgl")
    dgl.NID = "_ID"
    dgl.heterograph = lambda d: FakeHetero(*next(iter(d.values())))
    dgl.node_subgraph = _node_subgraph
    fn = types.ModuleType("dgl.function")
    fn.copy_u = lambda u, out: ("copy_u", u, out)
    fn.sum = lambda msg, out: ("sum", msg, out)
    fn.mean = lambda msg, out: ("mean", msg, out)
    dgl.function = fn
    data = types.ModuleType("dgl.data")
    data.RedditDataset = data.YelpDataset = object
    dist_m = types.ModuleType("dgl.distributed")
    dist_m.partition_graph = None
    nn_m = types.ModuleType("dgl.nn")
    nn_m.GATConv = GATConvStandIn
    dgl.data, dgl.distributed, dgl.nn = data, dist_m, nn_m
    ogb = types.ModuleType("ogb")
    ogbn = types.ModuleType("ogb.nodeproppred")
    ogbn.DglNodePropPredDataset = object
    for name, m in {"dgl": dgl, "dgl.function": fn, "dgl.data": data, "dgl.distributed": dist_m, "dgl.nn": nn_m,
                    "ogb": ogb, "ogb.nodeproppred": ogbn}.items():
        sys.modules[name] = m


# ------------------------------------------------------------------------------------------------------
# CUDA placement -> CPU
# ------------------------------------------------------------------------------------------------------
def patch_torch_for_cpu():
    def strip(kw):
        kw.pop("pin_memory", None)
        if "device" in kw and kw["device"] is not None and "cuda" in str(kw["device"]):
            kw["device"] = "cpu"
        return kw

    for name in ("zeros", "zeros_like", "as_tensor", "tensor", "empty"):
        orig = getattr(torch, name)
        setattr(torch, name, (lambda o: (lambda *a, **k: o(*a, **strip(k))))(orig))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    class _Stream:
        def __init__(self, *a, **k):
            pass

        def wait_stream(self, *_):
            pass

    torch.cuda.Stream = _Stream
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.reset_peak_memory_stats = lambda *a, **k: None
    torch.cuda.memory_allocated = torch.cuda.max_memory_allocated = torch.cuda.memory_reserved = lambda *a, **k: 0


class _LazyHandle:
    def __init__(self, fn):
        self.fn, self.done = fn, False

    def wait(self):
        if not self.done:
            self.done = True
            self.fn()


class _LazyPool:
    def __init__(self, processes=None):
        pass

    def apply_async(self, fn, args=()):
        return _LazyHandle(lambda: fn(*args))


# ------------------------------------------------------------------------------------------------------
def worker(rank, world, cfg, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, ROOT)
    import bns_gcn_b200  # noqa: F401  (only the data generator / partitioner: the INPUTS)
    from bns_gcn_b200.data import make_graph, partition_graph
    install_dgl_shim()
    patch_torch_for_cpu()
    sys.path.insert(0, REF)
    os.chdir(out_dir)                                   # run() creates checkpoint/ and results/ in the cwd
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helper.reducer as ref_reducer               # the reference's modules
    ref_reducer.ThreadPool = _LazyPool
    import train as ref_train
    import argparse

    fg = make_graph(cfg["shape"], seed=0, device=torch.device("cpu"), **cfg.get("graph_override", {}))
    part = partition_graph(fg, world, "random", seed=0, inductive=cfg.get("inductive", False),
                           device=torch.device("cpu"))[rank]
    lg = part.graph
    v = torch.repeat_interleave(torch.arange(lg.n_in), lg.indptr[1:] - lg.indptr[:-1])
    subg = FakeGraph(lg.indices.clone(), v, lg.n_in + lg.n_halo)
    node_dict = {k: t.clone() for k, t in part.node_dict.items()}

    class GPB:
        def partid2nids(self, i):
            return torch.arange(int(part.gpb.ranges[i]), int(part.gpb.ranges[i + 1]))

    args = argparse.Namespace(dataset=cfg.get("dataset", "synthetic"), model=cfg["model"], dropout=0.0, lr=1e-2, sampling_rate=cfg["rate"],
                              heads=cfg.get("heads", 1), n_epochs=cfg["epochs"], n_partitions=world, n_hidden=cfg["n_hidden"],
                              n_layers=cfg["n_layers"], log_every=1, weight_decay=0.0, norm=cfg.get("norm", "layer"),
                              n_linear=cfg.get("n_linear", 0), use_pp=True, inductive=cfg.get("inductive", False), seed=0, backend="gloo", eval=False,
                              graph_name="golden", n_feat=part.meta["n_feat"], n_class=part.meta["n_class"],
                              n_train=part.meta["n_train"])
    rec = {"selected": [], "logits": [], "layer_out": [], "loss": []}
    np.random.seed(1000 + rank)                         # the reference never seeds numpy (train.py:233)

    orig_select = ref_train.select_node

    def select_node(boundary, send_size):
        sel = orig_select(boundary, send_size)
        if cfg.get("philox_seed") is not None:
            # draw the sets with the product's counter-based sampler instead of numpy (same distribution): pins the
            # reference on exactly the index sets the CUDA path samples by itself
            from oracle import philox
            peers = [j for j in range(world) if j != rank]
            drawn = philox.sample_boundary([boundary[j].numpy() for j in peers], [int(send_size[j]) for j in peers],
                                           cfg["philox_seed"], len(rec["selected"]))
            sel = [None] * world
            for i, j in enumerate(peers):
                sel[j] = torch.from_numpy(np.ascontiguousarray(drawn[i])).long()
        rec["selected"].append([None if s is None else s.clone() for s in sel])
        return sel
    ref_train.select_node = select_node

    holder = {}
    orig_create = ref_train.create_model

    def create_model(layer_size, a):
        m = orig_create(layer_size, a)
        holder["model"] = m
        outs = {}
        for i, layer in enumerate(m.layers):
            layer.register_forward_hook(lambda mod, inp, out, i=i: outs.__setitem__(i, out.detach().clone()))

        def after(mod, inp, out):
            rec["logits"].append(out.detach().clone())
            rec["layer_out"].append([outs[i] for i in range(len(m.layers))])
        m.register_forward_hook(after)
        return m
    ref_train.create_model = create_model

    orig_pre = ref_train.precompute

    def precompute(*a, **k):
        f = orig_pre(*a, **k)
        rec["feat0"] = f.detach().clone()
        return f
    ref_train.precompute = precompute

    orig_boundary = ref_train.get_boundary

    def get_boundary(nd, gpb):
        b = orig_boundary(nd, gpb)
        rec["boundary"] = [None if x is None else x.clone() for x in b]
        return b
    ref_train.get_boundary = get_boundary

    ref_train.run(subg, node_dict, GPB(), args)        # <- the reference's own driver, unmodified

    m = holder["model"]
    if cfg.get("eval_logits") and rank == 0:
        # the reference's evaluation forward (train.py:44-49: model.eval(); model(g, feat)) on the WHOLE graph with the
        # trained weights: the layers' evaluation branches (degrees taken from the graph, module/layer.py:39-45, 93-102)
        full = FakeGraph(fg.src.clone(), fg.dst().clone(), fg.n_nodes)
        was_training = m.training
        m.eval()
        with torch.no_grad():
            rec["eval_logits"] = m(full, fg.feat.clone()).detach().clone()
        m.train(was_training)
    rec["params"] = [p.detach().clone() for p in m.parameters()]
    rec["grads"] = [p.grad.detach().clone() for p in m.parameters()]
    rec["param_names"] = [n for n, _ in m.named_parameters()]
    rec["config"] = dict(cfg)
    torch.save(rec, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def main():
    import tempfile
    import torch.multiprocessing as mp
    only = set(sys.argv[1:])                      # optional: names of the configs to (re)generate
    for i, (name, cfg) in enumerate(CONFIGS.items()):
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            mp.spawn(worker, args=(cfg["n_parts"], cfg, 29600 + i, d), nprocs=cfg["n_parts"], join=True)
            ranks = [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(cfg["n_parts"])]
        if cfg.get("slim"):
            keep0 = ("selected", "boundary", "param_names", "logits", "layer_out", "params", "grads", "eval_logits")
            ranks = [{k: ([v[-1]] if k in ("logits", "layer_out") else v) for k, v in rk.items()
                      if k in (keep0 if r == 0 else ("selected", "boundary", "param_names"))} for r, rk in enumerate(ranks)]
        out = os.path.join(HERE, f"ref_{name}_p{cfg['n_parts']}.pt")
        torch.save({"config": cfg, "ranks": ranks}, out)
        print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
