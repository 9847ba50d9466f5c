"""Command-line flags: the reference's flag set (helper/parser.py:4-61: same names, both ``--a-b`` and
``--a_b`` spellings, same defaults) plus the few that only exist here (marked NEW)."""
import argparse


def build_parser():
    parser = argparse.ArgumentParser(description='BNS-GCN (B200-native hot path)')
    parser.add_argument("--dataset", type=str, default='reddit',
                        help="synthetic shape to generate: reddit | ogbn-products | yelp | synthetic-10k | small | tiny")
    parser.add_argument("--data-path", "--data_path", type=str, default='./dataset/')
    parser.add_argument("--part-path", "--part_path", type=str, default='./partition/')
    parser.add_argument("--graph-name", "--graph_name", type=str, default='')
    parser.add_argument("--model", type=str, default='graphsage')
    parser.add_argument("--dropout", type=float, default=0.5)
    parser.add_argument("--lr", type=float, default=1e-2)
    parser.add_argument("--sampling-rate", "--sampling_rate", type=float, default=1)
    parser.add_argument("--heads", type=int, default=1)
    parser.add_argument("--n-epochs", "--n_epochs", type=int, default=200)
    parser.add_argument("--n-partitions", "--n_partitions", type=int, default=2)
    parser.add_argument("--n-hidden", "--n_hidden", type=int, default=16)
    parser.add_argument("--n-layers", "--n_layers", type=int, default=2)
    parser.add_argument("--log-every", "--log_every", type=int, default=10)
    parser.add_argument("--weight-decay", "--weight_decay", type=float, default=0)
    parser.add_argument("--norm", choices=['layer', 'batch'], default='layer')
    parser.add_argument("--partition-obj", "--partition_obj", choices=['vol', 'cut'], default='vol')
    parser.add_argument("--partition-method", "--partition_method", choices=['metis', 'random'], default='metis')
    parser.add_argument("--n-linear", "--n_linear", type=int, default=0)
    parser.add_argument("--use-pp", "--use_pp", action='store_true')
    parser.add_argument("--inductive", action='store_true')
    parser.add_argument("--fix-seed", "--fix_seed", action='store_true')
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--backend", type=str, default='nccl',
                        help="exchange transport: nccl (staged all-to-all) | p2p (peer-mapped slabs over NVLink); "
                             "the reference's gloo / mpi host-staged transports are what this replaces")
    parser.add_argument("--port", type=int, default=18118)
    parser.add_argument("--master-addr", "--master_addr", type=str, default="127.0.0.1")
    parser.add_argument("--node-rank", "--node_rank", type=int, default=0)
    parser.add_argument("--parts-per-node", "--parts_per_node", type=int, default=10)
    parser.add_argument('--skip-partition', action='store_true')
    parser.add_argument('--eval', action='store_true')
    parser.add_argument('--no-eval', action='store_false', dest='eval')
    parser.add_argument("--sampler-seed", "--sampler_seed", type=int, default=0,
                        help="NEW: Philox seed of the boundary sampler (the reference draws from unseeded numpy)")
    parser.set_defaults(eval=True)
    return parser


def create_parser(argv=None):
    return build_parser().parse_args(argv)
