"""Command-line flags.  The reference's flag set (helper/parser.py:4-61) is an interface, so names, spellings
(``--a-b`` and ``--a_b``), types and defaults are kept; it is declared as a table here, followed by the flags that
only exist in this build."""
import argparse

# (flag, type or None for a switch, default, extra argparse keywords)
_REFERENCE_FLAGS = [
    ("dataset", str, "reddit", dict(help="synthetic shape to generate: reddit | ogbn-products | yelp | synthetic-10k | "
                                         "small | tiny")),
    ("data-path", str, "./dataset/", {}),
    ("part-path", str, "./partition/", {}),
    ("graph-name", str, "", {}),
    ("model", str, "graphsage", {}),
    ("dropout", float, 0.5, {}),
    ("lr", float, 1e-2, {}),
    ("sampling-rate", float, 1, {}),
    ("heads", int, 1, {}),
    ("n-epochs", int, 200, {}),
    ("n-partitions", int, 2, {}),
    ("n-hidden", int, 16, {}),
    ("n-layers", int, 2, {}),
    ("log-every", int, 10, {}),
    ("weight-decay", float, 0, {}),
    ("norm", None, "layer", dict(choices=["layer", "batch"])),
    ("partition-obj", None, "vol", dict(choices=["vol", "cut"])),
    ("partition-method", None, "metis", dict(choices=["metis", "random"])),
    ("n-linear", int, 0, {}),
    ("use-pp", "switch", False, {}),
    ("inductive", "switch", False, {}),
    ("fix-seed", "switch", False, {}),
    ("seed", int, 0, {}),
    ("backend", str, "nccl", dict(help="exchange transport: nccl (staged all-to-all) | p2p (peer-mapped slabs over "
                                       "NVLink); the reference's gloo / mpi host-staged transports are what this "
                                       "replaces")),
    ("port", int, 18118, {}),
    ("master-addr", str, "127.0.0.1", {}),
    ("node-rank", int, 0, {}),
    ("parts-per-node", int, 10, {}),
]


def _spellings(flag: str):
    dashed = "--" + flag
    return (dashed,) if "-" not in flag else (dashed, dashed.replace("-", "_").replace("__", "--", 1))


def build_parser():
    parser = argparse.ArgumentParser(description='BNS-GCN (B200-native hot path)')
    for flag, kind, default, extra in _REFERENCE_FLAGS:
        if kind == "switch":
            parser.add_argument(*_spellings(flag), action='store_true')
        elif kind is None:
            parser.add_argument(*_spellings(flag), default=default, **extra)
        else:
            parser.add_argument(*_spellings(flag), type=kind, default=default, **extra)
    parser.add_argument('--skip-partition', action='store_true')
    # --eval / --no-eval write the same destination; evaluation is on unless --no-eval is given (parser.py:57-59)
    parser.add_argument('--eval', action='store_true')
    parser.add_argument('--no-eval', action='store_false', dest='eval')
    parser.set_defaults(eval=True)
    # only here
    parser.add_argument(*_spellings("sampler-seed"), type=int, default=0,
                        help="NEW: Philox seed of the boundary sampler (the reference draws from unseeded numpy)")
    return parser


def create_parser(argv=None):
    return build_parser().parse_args(argv)
