"""Exchange metadata and p2p helpers of the reference's helper/utils.py, on the partition contract of
``data/partition.py`` instead of DGL objects.  Host-side integer logic (works on CPU tensors under gloo, which is
how the world_size-2 CPU tests drive it, and on CUDA tensors under NCCL / in-process ranks)."""
from __future__ import annotations

from typing import List, Optional

import torch

from ..data.partition import NID
from . import context as ctx


class TransferTag:            # helper/utils.py:15-18
    NODE = 0
    FEAT = 1
    DEG = 2


def get_layer_size(n_feat, n_hidden, n_class, n_layers):
    """helper/utils.py:143-147."""
    layer_size = [n_feat]
    layer_size.extend([n_hidden] * (n_layers - 1))
    layer_size.append(n_class)
    return layer_size


def get_boundary(node_dict, gpb) -> List[Optional[torch.Tensor]]:
    """helper/utils.py:150-184.  ``boundary[j]``: sorted local ids of MY inner nodes that are halo nodes of
    rank ``j`` (what I may be asked to send to ``j``).  Rank r tells each peer which of the peer's nodes it holds
    as halo; the peer sorts the list."""
    c = ctx.comm()
    rank, size = c.rank, c.size
    boundary: List[Optional[torch.Tensor]] = [None] * size
    if size == 1:
        return boundary
    dev = node_dict["part_id"].device
    counts_out = [None] * size
    lists_out = [None] * size
    for j in range(size):
        if j == rank:
            continue
        belong = node_dict["part_id"] == j
        lists_out[j] = (node_dict[NID][belong] - int(gpb.ranges[j])).contiguous()
        counts_out[j] = torch.tensor([lists_out[j].numel()], dtype=torch.int64, device=dev)
    counts_in = [None if j == rank else torch.zeros(1, dtype=torch.int64, device=dev) for j in range(size)]
    c.alltoall(counts_out, counts_in, tag=128)
    lists_in = [None if j == rank else torch.zeros(int(counts_in[j]), dtype=torch.int64, device=dev)
                for j in range(size)]
    c.alltoall(lists_out, lists_in, tag=129)
    for j in range(size):
        if j != rank:
            boundary[j] = torch.sort(lists_in[j])[0]
    return boundary


def data_transfer(data, recv_shape, tag, dtype=torch.float):
    """helper/utils.py:190-213: every rank sends ``data[j]`` to ``j`` and returns what the others sent it.
    The reference stages through pinned host buffers and a gloo ring; here the tensors stay on the device."""
    c = ctx.comm()
    rank, size = c.rank, c.size
    res: List[Optional[torch.Tensor]] = [None] * size
    if size == 1:
        return res
    dev = next(d for d in data if d is not None).device
    send = [None if j == rank else data[j].to(dtype).contiguous() for j in range(size)]
    for j in range(size):
        if j != rank:
            res[j] = torch.zeros(tuple(recv_shape[j]), dtype=dtype, device=dev)
    c.alltoall(send, res, tag=tag)
    return res


def merge_feature(feat, recv):
    """helper/utils.py:216-223: ``[feat | recv_0 | recv_1 ...]`` in rank order, own rank skipped."""
    return torch.cat([feat] + [r for r in recv if r is not None])


def minus_one_tensor(size, device=None):
    return torch.full((size,), -1, dtype=torch.long, device=device)


def nonzero_idx(x):
    return torch.nonzero(x, as_tuple=True)[0]


def print_memory(s):
    """helper/utils.py:244-250."""
    rank = ctx.comm().rank
    torch.cuda.synchronize()
    print('(rank %d) ' % rank + s + ': current {:.2f}MB, peak {:.2f}MB, reserved {:.2f}MB'.format(
        torch.cuda.memory_allocated() / 1024 / 1024, torch.cuda.max_memory_allocated() / 1024 / 1024,
        torch.cuda.memory_reserved() / 1024 / 1024))
