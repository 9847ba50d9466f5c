"""Host-side pieces of SURVEY §8(f) ranks 3-4: the on-disk partition store (graph_partition / load_partition of
helper/utils.py:73-140) and the evaluation / checkpoint helpers (train.py:14-61, 427-456).  CPU only."""
import argparse
import json
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _args(tmp, **kw):
    d = dict(dataset="tiny", n_partitions=3, partition_method="random", partition_obj="vol", inductive=False,
             part_path=str(tmp), graph_name="", sampling_rate=0.5)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("inductive", [False, True])
def test_store_round_trip_equals_in_memory_partition(tmp_path, inductive):
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.data import make_graph, partition_graph, graph_partition, load_partition, default_graph_name
    cpu = torch.device("cpu")
    a = _args(tmp_path, inductive=inductive)
    fg = make_graph("tiny", seed=0, device=cpu)
    cfg_path = graph_partition(a, fg=fg, device=cpu)
    assert a.graph_name == default_graph_name(a) == "tiny-3-random-vol-" + ("induc" if inductive else "trans")
    assert os.path.basename(cfg_path) == a.graph_name + ".json"
    want = partition_graph(fg, 3, "random", seed=0, inductive=inductive, device=cpu)
    with open(os.path.join(os.path.dirname(cfg_path), "meta.json")) as f:
        meta = json.load(f)
    assert meta == {"n_feat": fg.n_feat, "n_class": fg.n_class, "n_train": int(fg.train_mask.sum())}
    for r in range(3):
        b = _args(tmp_path, inductive=inductive)
        g, nd, gpb = load_partition(b, r)
        w = want[r]
        assert (b.n_feat, b.n_class, b.n_train) == (w.meta["n_feat"], w.meta["n_class"], w.meta["n_train"])
        assert (g.n_in, g.n_halo) == (w.graph.n_in, w.graph.n_halo)
        assert g.indptr.dtype == torch.int64 and g.indices.dtype == torch.int64
        assert torch.equal(g.indptr, w.graph.indptr) and torch.equal(g.indices, w.graph.indices)
        assert torch.equal(gpb.ranges, w.gpb.ranges)
        assert set(nd) == set(w.node_dict)
        assert ("val_mask" in nd) == (not inductive)
        for k, v in w.node_dict.items():
            assert nd[k].dtype == v.dtype, k                       # bool masks come back as bool (utils.py:114-128)
            assert torch.equal(nd[k], v), k


def test_store_partitions_once_and_checks_its_inputs(tmp_path):
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.data import make_graph, graph_partition, load_partition
    cpu = torch.device("cpu")
    a = _args(tmp_path)
    with pytest.raises(FileNotFoundError):
        load_partition(_args(tmp_path), 0)
    fg = make_graph("tiny", seed=0, device=cpu)
    cfg_path = graph_partition(a, fg=fg, device=cpu)
    first = os.path.getmtime(os.path.join(os.path.dirname(cfg_path), "part0", "indices.npy"))
    os.remove(os.path.join(os.path.dirname(cfg_path), "meta.json"))
    graph_partition(a, fg=fg, device=cpu)                          # utils.py:86: parts are kept, meta.json rewritten
    assert os.path.getmtime(os.path.join(os.path.dirname(cfg_path), "part0", "indices.npy")) == first
    assert os.path.exists(os.path.join(os.path.dirname(cfg_path), "meta.json"))
    with pytest.raises(RuntimeError):
        load_partition(_args(tmp_path, n_partitions=4, graph_name=a.graph_name), 0)
    with pytest.raises(IndexError):
        load_partition(_args(tmp_path), 3)


def test_calc_acc_matches_sklearn_micro_f1():
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.evaluate import calc_acc
    from sklearn.metrics import f1_score
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(200, 7, generator=g)
    labels = torch.randint(0, 7, (200,), generator=g)
    assert calc_acc(logits, labels) == (logits.argmax(1) == labels).sum().item() / 200
    ml = (torch.rand(200, 7, generator=g) < 0.3).float()
    assert abs(calc_acc(logits, ml) - f1_score(ml.numpy(), (logits > 0).numpy(), average="micro")) < 1e-12
    assert calc_acc(torch.full((4, 3), -1.0), torch.zeros(4, 3)) == 0.0      # no positives anywhere


def test_checkpoint_keys_are_the_references_and_round_trip(tmp_path):
    """State-dict keys of the model mirrors == the parameter names the reference's own model reported when the golden
    vectors were minted; save / load restores every tensor bit for bit."""
    import torch.nn.functional as F
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.evaluate import checkpoint_path, load_checkpoint, result_file_name, save_checkpoint
    from bns_gcn_b200.module.model import GraphSAGE
    gold = torch.load(os.path.join(GOLD, "ref_graphsage_p2.pt"))
    cfg, r0 = gold["config"], gold["ranks"][0]
    n_feat, n_class = r0["feat0"].shape[1] // 2, r0["logits"][-1].shape[1]
    layer_size = [n_feat] + [cfg["n_hidden"]] * (cfg["n_layers"] - 1) + [n_class]
    torch.manual_seed(0)
    m = GraphSAGE(layer_size, F.relu, use_pp=True, dropout=0.0, norm="layer")
    assert [k for k, _ in m.named_parameters()] == r0["param_names"]
    assert list(m.state_dict().keys()) == r0["param_names"]
    assert [tuple(p.shape) for p in m.parameters()] == [tuple(p.shape) for p in r0["params"]]
    a = argparse.Namespace(graph_name="tiny-2-random-vol-trans", sampling_rate=0.5, dataset="tiny", n_partitions=2)
    assert checkpoint_path(a, 9) == "checkpoint/tiny-2-random-vol-trans_p0.50_9.pth.tar"      # train.py:428
    assert checkpoint_path(a) == "checkpoint/tiny-2-random-vol-trans_final.pth.tar"            # train.py:452
    assert result_file_name(a) == "results/tiny_n2_p0.50.txt"                                  # train.py:356
    path = os.path.join(tmp_path, checkpoint_path(a, 9))
    save_checkpoint(m, path)
    torch.manual_seed(1)
    m2 = GraphSAGE(layer_size, F.relu, use_pp=True, dropout=0.0, norm="layer")
    load_checkpoint(m2, path)
    for (k, v), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k == k2 and torch.equal(v, v2)
    # a checkpoint made of the reference's own tensors loads too
    torch.save({k: v for k, v in zip(r0["param_names"], r0["params"])}, os.path.join(tmp_path, "ref.pth.tar"))
    load_checkpoint(m2, os.path.join(tmp_path, "ref.pth.tar"))
    assert all(torch.equal(p, q) for p, q in zip(m2.parameters(), r0["params"]))


@pytest.mark.parametrize("kind", ["graphsage", "gcn", "gat"])
def test_model_mirrors_start_from_the_references_weights(kind):
    """Under ``torch.manual_seed(seed)`` (train.py:331-333) the reference's initial weights are a function of the order
    in which sub-modules are created and re-initialised.  The oracle restates that order (and is pinned to the
    reference by the golden vectors); the product's mirrors must reproduce it: same names, same tensors."""
    import torch.nn.functional as F
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.module.model import GAT, GCN, GraphSAGE
    from oracle import bns_oracle as O
    for use_pp in (False, True):
        for norm in ("layer", "batch"):
            for n_linear in (0, 1):
                if kind == "gat" and not use_pp:
                    continue                                   # the reference's GAT path requires --use-pp
                layer_size = [12, 16, 16, 5]
                torch.manual_seed(4)
                if kind == "gat":
                    ours = GAT(layer_size, F.relu, use_pp, heads=2, dropout=0.1, norm=norm, train_size=33, n_linear=n_linear)
                else:
                    ours = (GraphSAGE if kind == "graphsage" else GCN)(layer_size, F.relu, use_pp, dropout=0.1, norm=norm,
                                                                         train_size=33, n_linear=n_linear)
                torch.manual_seed(4)
                ref = O.build_model(kind, layer_size, use_pp, 0.1, norm, 33, n_linear, 2)
                a, b = dict(ours.named_parameters()), dict(ref.named_parameters())
                assert list(a) == list(b), (use_pp, norm, n_linear)
                for k in a:
                    assert torch.equal(a[k], b[k]), (k, use_pp, norm, n_linear)


@pytest.mark.parametrize("kind", ["graphsage", "gcn"])
def test_all_linear_stack_runs_on_the_cpu_and_matches_the_oracle(kind):
    """With ``n_linear == n_layers`` the stack holds no graph layer, so the product's ``_forward`` / inter-layer step
    run on plain ATen CPU ops: dropout -> Linear -> norm -> activation must equal the oracle's (reference's) loop,
    in training mode too (same dropout draws)."""
    import torch.nn.functional as F
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.module.model import GCN, GraphSAGE
    from oracle import bns_oracle as O
    layer_size = [12, 16, 16, 5]
    for p in (0.0, 0.4):
        torch.manual_seed(4)
        ours = (GraphSAGE if kind == "graphsage" else GCN)(layer_size, F.relu, False, dropout=p, norm="layer", n_linear=3)
        torch.manual_seed(4)
        ref = O.build_model(kind, layer_size, False, p, "layer", None, 3, 1)
        x = torch.randn(40, 12)
        for training in (False, True):
            ours.train(training)
            ref.train(training)
            torch.manual_seed(8)
            a = ours(None, x)
            torch.manual_seed(8)
            b = ref(None, x)
            assert torch.equal(a, b), (p, training)
