"""Launcher with the reference's flags (main.py:10-64): generate (instead of download) the graph, partition it,
start one process per partition / GPU and run ``train.run`` in each.

    python -m bns_gcn_b200.main --dataset reddit --n-partitions 4 --model graphsage --n-layers 3 --n-hidden 256 \
        --sampling-rate 0.1 --use-pp --partition-method random --n-epochs 50 --no-eval

Under torchrun (RANK / WORLD_SIZE set) it joins the existing job instead of spawning.
"""
import os
import random
import warnings

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from .data import make_graph, partition_graph
from .helper.parser import create_parser


def init_processes(rank, size, args):
    """train.py:459-470: rendezvous, load this rank's partition, run."""
    from . import train
    os.environ.setdefault('MASTER_ADDR', args.master_addr)
    os.environ.setdefault('MASTER_PORT', '%d' % args.port)
    local = rank % max(torch.cuda.device_count(), 1)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', rank=rank, world_size=size, device_id=dev)
    fg = make_graph(args.dataset, seed=0, device=dev)
    part = partition_graph(fg, size, args.partition_method, seed=0, inductive=args.inductive, ranks=[rank],
                           device=dev)[0]
    del fg
    args.n_feat, args.n_class, args.n_train = part.meta['n_feat'], part.meta['n_class'], part.meta['n_train']
    if args.eval:
        warnings.warn('evaluation / checkpointing (train.py:427-456) is not on the rebuilt path; running --no-eval')
        args.eval = False
    train.run(part.graph, part.node_dict, part.gpb, args, dev)
    dist.destroy_process_group()


def main(argv=None):
    args = create_parser(argv)
    if args.fix_seed is False:
        if args.parts_per_node < args.n_partitions:
            warnings.warn('Please enable `--fix-seed` for multi-node training.')
        args.seed = random.randint(0, 1 << 31)
    if args.backend in ('gloo', 'mpi'):
        warnings.warn(f'--backend {args.backend}: host-staged transports are what this build replaces; using nccl')
        args.backend = 'nccl'
    print(args)
    if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ:
        init_processes(int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), args)
        return
    mp.set_start_method('spawn', force=True)
    start = args.node_rank * args.parts_per_node
    procs = []
    for i in range(start, min(start + args.parts_per_node, args.n_partitions)):
        p = mp.Process(target=init_processes, args=(i, args.n_partitions, args))
        p.start()
        procs.append(p)
    for p in procs:
        p.join()


if __name__ == '__main__':
    main()
