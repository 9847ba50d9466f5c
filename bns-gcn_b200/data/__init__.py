from .synthetic import FullGraph, SHAPES, make_graph
from .partition import (NID, GraphPartitionBook, LocalGraph, Partition, partition_graph, extract_partition,
                        assign_parts, relabel, induced_subgraph)
