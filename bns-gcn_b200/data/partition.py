"""Vertex partitioner that emits the data contract of the reference's partition loader.

The reference calls ``dgl.distributed.partition_graph`` / ``load_partition``
(``helper/utils.py:73-140``); DGL and METIS are not available here, so this file
produces, per rank, exactly what ``load_partition`` returns and ``train.run``
relies on:

* node ids are relabelled so that partition ``i`` owns the contiguous global range
  ``[ranges[i], ranges[i+1])``  (``gpb.partid2nids``, used at ``train.py:97-101`` and
  ``utils.py:167-168``);
* the local graph holds **all in-edges of the inner nodes**; inner nodes have local ids
  ``[0, n_in)``, the 1-hop halo nodes follow (``train.py:85-86`` depends on this);
* ``node_dict['_ID']`` (global id), ``['part_id']`` (owner) and ``['inner_node']`` have one entry
  per local node; ``feat/label/in_deg/out_deg/train_mask(/val_mask/test_mask)`` one entry per
  inner node; degrees are those of the *full* graph (``utils.py:92-93``);
* ``meta`` = ``n_feat, n_class, n_train`` with the global train count (``utils.py:97-98``).

``--partition-method random`` is the reference's own option (``helper/parser.py:37``).
``metis`` is served by a stand-in (reverse Cuthill-McKee order cut into equal blocks):
the objective differs from METIS but the contract is the same.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np
import torch

from .synthetic import FullGraph

NID = "_ID"   # the key DGL uses for ``dgl.NID``


@dataclasses.dataclass
class GraphPartitionBook:
    """Stand-in for DGL's ``gpb``: only ``partid2nids`` is used by the path (``train.py:97-98``)."""
    ranges: torch.Tensor   # int64 [P+1]

    def partid2nids(self, i: int) -> torch.Tensor:
        return torch.arange(int(self.ranges[i]), int(self.ranges[i + 1]), dtype=torch.int64)

    def num_partitions(self) -> int:
        return int(self.ranges.numel() - 1)


@dataclasses.dataclass
class LocalGraph:
    """The ``subg`` of ``load_partition``: in-edges of inner nodes, CSR by (inner) destination.

    ``indices`` are local source ids: ``< n_in`` inner, ``>= n_in`` halo.
    """
    n_in: int
    n_halo: int
    indptr: torch.Tensor    # int64 [n_in+1]
    indices: torch.Tensor   # int64 [E_local]

    def num_nodes(self) -> int:
        return self.n_in + self.n_halo

    def num_edges(self) -> int:
        return int(self.indices.numel())


@dataclasses.dataclass
class Partition:
    rank: int
    n_parts: int
    graph: LocalGraph
    node_dict: Dict[str, torch.Tensor]
    gpb: GraphPartitionBook
    meta: Dict[str, int]


def assign_parts(fg: FullGraph, n_parts: int, method: str, seed: int) -> torch.Tensor:
    """Owner of every node, int64 ``[N]``, balanced to ±1 node."""
    n = fg.n_nodes
    if n_parts == 1:
        return torch.zeros(n, dtype=torch.int64)
    if method == "random":
        gen = torch.Generator().manual_seed(seed + 7919)
        order = torch.randperm(n, generator=gen)
    elif method == "metis":
        import scipy.sparse as sp
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        a = sp.csr_matrix((np.ones(fg.n_edges, dtype=np.int8), fg.src.numpy(), fg.indptr.numpy()), shape=(n, n))
        order = torch.from_numpy(np.ascontiguousarray(reverse_cuthill_mckee(a, symmetric_mode=True)).astype(np.int64))
    else:
        raise ValueError(f"unknown partition method {method!r}")
    part = torch.empty(n, dtype=torch.int64)
    part[order] = (torch.arange(n, dtype=torch.int64) * n_parts) // n
    return part


def relabel(fg: FullGraph, part: torch.Tensor, n_parts: int, device: Optional[torch.device] = None):
    """Renumber nodes so every partition is a contiguous id range (what DGL's partitioner does)."""
    if device is None:
        device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    n = fg.n_nodes
    order = torch.argsort(part, stable=True)              # new id -> old id
    new_id = torch.empty(n, dtype=torch.int64)
    new_id[order] = torch.arange(n, dtype=torch.int64)
    counts = torch.bincount(part, minlength=n_parts)
    ranges = torch.zeros(n_parts + 1, dtype=torch.int64)
    ranges[1:] = torch.cumsum(counts, 0)
    nid = new_id.to(device)
    dst = nid[fg.dst().to(device)]
    src = nid[fg.src.to(device)]
    perm = torch.argsort(dst * n + src)
    dst, src = dst[perm], src[perm]
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0)
    g = FullGraph(n, indptr.cpu(), src.cpu(), fg.feat[order], fg.label[order], fg.train_mask[order],
                  fg.val_mask[order], fg.test_mask[order], fg.n_class)
    return g, ranges


def induced_subgraph(fg: FullGraph, mask: torch.Tensor) -> FullGraph:
    """``g.subgraph(mask)`` (``utils.py:77`` inductive setting): keep edges with both ends in ``mask``."""
    n = fg.n_nodes
    keep = torch.nonzero(mask, as_tuple=True)[0]
    new_id = torch.full((n,), -1, dtype=torch.int64)
    new_id[keep] = torch.arange(keep.numel(), dtype=torch.int64)
    dst = new_id[fg.dst()]
    src = new_id[fg.src]
    ok = (dst >= 0) & (src >= 0)
    dst, src = dst[ok], src[ok]                            # order (dst, src) is preserved
    m = keep.numel()
    indptr = torch.zeros(m + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(torch.bincount(dst, minlength=m), 0)
    return FullGraph(m, indptr, src, fg.feat[keep], fg.label[keep], fg.train_mask[keep],
                     fg.val_mask[keep], fg.test_mask[keep], fg.n_class)


def extract_partition(g: FullGraph, ranges: torch.Tensor, rank: int, inductive: bool = False,
                      in_deg: Optional[torch.Tensor] = None, out_deg: Optional[torch.Tensor] = None) -> Partition:
    """Cut rank ``rank``'s piece out of a relabelled graph (``load_partition``, ``utils.py:101-140``)."""
    n_parts = int(ranges.numel() - 1)
    start, end = int(ranges[rank]), int(ranges[rank + 1])
    n_in = end - start
    e0, e1 = int(g.indptr[start]), int(g.indptr[end])
    src = g.src[e0:e1]
    inner = (src >= start) & (src < end)
    halo = torch.unique(src[~inner])                       # sorted global ids
    local = torch.where(inner, src - start, n_in + torch.searchsorted(halo, src))
    indptr = (g.indptr[start:end + 1] - e0).clone()
    if in_deg is None:
        in_deg = g.in_degrees()
    if out_deg is None:
        out_deg = g.out_degrees()
    gid = torch.cat([torch.arange(start, end, dtype=torch.int64), halo])
    part_id = torch.searchsorted(ranges, gid, right=True) - 1
    inner_node = torch.zeros(gid.numel(), dtype=torch.bool)
    inner_node[:n_in] = True
    nd = {
        NID: gid,
        "part_id": part_id,
        "inner_node": inner_node,
        "feat": g.feat[start:end].clone(),
        "label": g.label[start:end].clone(),
        "in_deg": in_deg[start:end].clone(),
        "out_deg": out_deg[start:end].clone(),
        "train_mask": g.train_mask[start:end].clone(),
    }
    if not inductive:
        nd["val_mask"] = g.val_mask[start:end].clone()
        nd["test_mask"] = g.test_mask[start:end].clone()
    meta = {"n_feat": g.n_feat, "n_class": g.n_class, "n_train": int(g.train_mask.sum())}
    return Partition(rank, n_parts, LocalGraph(n_in, int(halo.numel()), indptr, local.contiguous()), nd,
                     GraphPartitionBook(ranges.clone()), meta)


def partition_graph(fg: FullGraph, n_parts: int, method: str = "random", seed: int = 0,
                    inductive: bool = False, ranks: Optional[List[int]] = None,
                    device: Optional[torch.device] = None) -> List[Partition]:
    """``graph_partition`` + ``load_partition`` in one call; returns the pieces for ``ranks`` (default all)."""
    if inductive:
        fg = induced_subgraph(fg, fg.train_mask)
    part = assign_parts(fg, n_parts, method, seed)
    g, ranges = relabel(fg, part, n_parts, device)
    in_deg, out_deg = g.in_degrees(), g.out_degrees()
    if ranks is None:
        ranks = list(range(n_parts))
    return [extract_partition(g, ranges, r, inductive, in_deg, out_deg) for r in ranks]
