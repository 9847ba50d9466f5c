"""CPU tests of the oracle (test infrastructure) -- these pin it:
 * against the golden vectors minted by the reference's own code (tests/golden/make_golden.py),
 * against the known-answer properties of SURVEY.md §4 (P-invariance, SpMM vs an independent index_add_),
 * Philox4x32-10 against the published Random123 known-answer vectors."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30)).item()


def _oracle_run(cfg, selected_per_epoch):
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.data import make_graph, partition_graph
    from oracle import bns_oracle as O
    fg = make_graph(cfg["shape"], seed=0, device=torch.device("cpu"), **cfg.get("graph_override", {}))
    parts = partition_graph(fg, cfg["n_parts"], "random", seed=0, inductive=cfg.get("inductive", False),
                            device=torch.device("cpu"))

    def fn(comm, r):
        rk = O.OracleRank(O.RankInput.from_partition(parts[r]), comm, model=cfg["model"], n_layers=cfg["n_layers"],
                          n_hidden=cfg["n_hidden"], sampling_rate=cfg["rate"], dropout=0.0, seed=0,
                          norm=cfg.get("norm", "layer"), n_linear=cfg.get("n_linear", 0), heads=cfg.get("heads", 1),
                          multilabel=(cfg.get("dataset") == "yelp"))
        for e in range(cfg["epochs"]):
            rk.epoch(selected=selected_per_epoch[e][r], trace=True)
        return rk
    return O.run_threads(cfg["n_parts"], fn)


@pytest.mark.parametrize("name", ["graphsage", "gcn", "graphsage_bn", "gat", "gat_yelp"])
def test_oracle_reproduces_reference_golden(name):
    """The oracle, fed the index sets the reference drew, reproduces what the reference computed:
    boundary sets exactly; precomputed features, layer outputs, logits, reduced grads, updated weights to 1e-5."""
    gold = torch.load(os.path.join(GOLD, f"ref_{name}_p2.pt"))
    cfg, ranks = gold["config"], gold["ranks"]
    sel = [[ranks[r]["selected"][e] for r in range(cfg["n_parts"])] for e in range(cfg["epochs"])]
    out = _oracle_run(cfg, sel)
    # --norm batch: a bias that feeds a BatchNorm has an exactly-zero true gradient (BN removes the mean), so what
    # reaches Adam is rounding noise that it normalises into +-lr steps; BN cancels the resulting shift downstream.
    # Those biases, and the pre-BN layer outputs they shift, are not comparable; everything else is.
    bn = cfg.get("norm") == "batch"
    last = cfg["n_layers"] - 1
    for r, rk in enumerate(out):
        g = ranks[r]
        for j, b in enumerate(g["boundary"]):
            if b is not None:
                assert torch.equal(rk.boundary[j], b)
        assert _rel(rk.feat, g["feat0"]) < 1e-6
        for i, lo in enumerate(g["layer_out"][-1]):
            if bn and i < last:
                continue
            if lo.dim() == 3:                      # GATConv returns [n, heads, F]; the model (and the trace) averages heads
                lo = lo.mean(1)
            assert _rel(rk.trace[f"layer{i}"], lo) < 1e-5, (r, i)
        assert _rel(rk.trace["logits"], g["logits"][-1]) < 1e-5
        for k, (p, gp, gg) in enumerate(zip(rk.net.parameters(), g["params"], g["grads"])):
            nm = g["param_names"][k]
            if bn and nm.endswith("bias") and nm.startswith("layers.") and int(nm.split(".")[1]) < last:
                continue
            assert _rel(p.detach(), gp) < 1e-5, (r, nm)
            assert _rel(p.grad, gg) < 1e-5, (r, nm)
        # parameter order / names are the reference's
        assert [n for n, _ in rk.net.named_parameters()] == g["param_names"]


@pytest.mark.parametrize("model", ["graphsage", "gcn"])
def test_oracle_p_invariance(model):
    """Sampling rate 1, dropout 0: any partitioning gives the single-partition loss and gradients."""
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.data import make_graph, partition_graph
    from oracle import bns_oracle as O
    fg = make_graph("tiny", seed=0, device=torch.device("cpu"))
    ref = None
    for P in (1, 2, 3):
        parts = partition_graph(fg, P, "random", seed=0, device=torch.device("cpu"))

        def fn(comm, r):
            rk = O.OracleRank(O.RankInput.from_partition(parts[r]), comm, model=model, n_layers=3, n_hidden=16,
                              sampling_rate=1.0, dropout=0.0, seed=0)
            loss = [rk.epoch(rng=np.random.RandomState(r)) for _ in range(2)]
            return loss, [p.grad.clone() for p in rk.net.parameters()]
        res = O.run_threads(P, fn)
        loss = [sum(x[0][e] for x in res) for e in range(2)]
        if ref is None:
            ref = (loss, res[0][1])
            continue
        for a, b in zip(loss, ref[0]):
            assert abs(a - b) < 1e-5 * abs(b)
        for a, b in zip(res[0][1], ref[1]):
            assert _rel(a, b) < 1e-5


def test_c_spmm_matches_index_add():
    from oracle import bns_oracle as O
    g = torch.Generator().manual_seed(0)
    u = torch.randint(0, 500, (20000,), generator=g)
    v = torch.randint(0, 300, (20000,), generator=g)
    x = torch.randn(500, 37, generator=g, requires_grad=True)
    e = O.EdgeList(u, v, 500, 300)
    y = O.CopyUSum.apply(e, x)
    assert _rel(y.detach(), O.copy_u_sum_indexadd(e, x.detach())) < 1e-6
    w = torch.randn(300, 37, generator=g)
    (y * w).sum().backward()
    ref = torch.zeros(500, 37).index_add_(0, u, w[v])
    assert _rel(x.grad, ref) < 1e-6
    # empty rows / empty graph
    e0 = O.EdgeList(torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.int64), 5, 4)
    assert torch.all(O.CopyUSum.apply(e0, torch.randn(5, 3)) == 0)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    from oracle.philox import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(x[0]) for x in got) == want


def test_philox_sampler_properties():
    from oracle.philox import sample_boundary
    b = [np.arange(0, 3000, 3, dtype=np.int64), np.arange(5, 905, dtype=np.int64), np.empty(0, dtype=np.int64)]
    k = [100, 900, 0]
    s = sample_boundary(b, k, seed=7, offset=3)
    for bi, ki, si in zip(b, k, s):
        assert len(si) == ki and len(np.unique(si)) == ki and np.isin(si, bi).all()
    assert np.array_equal(np.sort(s[1]), b[1])                   # k == b: a permutation
    s2 = sample_boundary(b, k, seed=7, offset=4)
    assert not np.array_equal(s[0], s2[0])
    assert np.array_equal(s[0], sample_boundary(b, k, seed=7, offset=3)[0])


@pytest.mark.parametrize("fixture", ["graphsage_small_p8", "gcn_small_p4", "graphsage_nlin_induc_p3", "graphsage_rate0_p2"])
def test_oracle_reproduces_reference_golden_slim(fixture):
    """The same pin on wider configurations (slim fixtures: every rank's index sets, rank 0's tensors): 8 partitions
    (7 peers per rank, 6000-node graph); GCN over 4 ranks at sampling rate 0.1; GraphSAGE with a trailing nn.Linear
    (--n-linear 1) on the inductive graph over 3 ranks; --sampling-rate 0 (nothing is ever exchanged, ratio 0).  The
    reference's own train.run on gloo vs the oracle fed the index sets the reference drew."""
    gold = torch.load(os.path.join(GOLD, f"ref_{fixture}.pt"))
    cfg, ranks = gold["config"], gold["ranks"]
    sel = [[ranks[r]["selected"][e] for r in range(cfg["n_parts"])] for e in range(cfg["epochs"])]
    out = _oracle_run(cfg, sel)
    for r, rk in enumerate(out):
        for j, b in enumerate(ranks[r]["boundary"]):
            if b is not None:
                assert torch.equal(rk.boundary[j], b)
    g0 = ranks[0]
    for i, lo in enumerate(g0["layer_out"][-1]):
        assert _rel(out[0].trace[f"layer{i}"], lo) < 1e-5, i
    assert _rel(out[0].trace["logits"], g0["logits"][-1]) < 1e-5
    for r, rk in enumerate(out):
        for p, gp, gg, nm in zip(rk.net.parameters(), g0["params"], g0["grads"], g0["param_names"]):
            assert _rel(p.detach(), gp) < 1e-5, (r, nm)
            assert _rel(p.grad, gg) < 1e-5, (r, nm)


@pytest.mark.parametrize("model", ["graphsage", "gcn"])
def test_oracle_evaluation_forward_reproduces_the_references(model):
    """tests/golden/ref_<model>_eval_p2.pt: after two training epochs make_golden.py runs the reference's evaluation
    forward (train.py:44-49: model.eval(); model(g, feat)) on the whole graph.  The oracle, trained on the same index
    sets, evaluated the same way (layers' evaluation branches, module/layer.py:39-45, 93-102: degrees from the graph,
    layer 0 concatenates [feat | mean] itself) gives the same logits."""
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.data import make_graph
    from oracle import bns_oracle as O
    gold = torch.load(os.path.join(GOLD, f"ref_{model}_eval_p2.pt"))
    cfg, ranks = gold["config"], gold["ranks"]
    sel = [[ranks[r]["selected"][e] for r in range(cfg["n_parts"])] for e in range(cfg["epochs"])]
    out = _oracle_run(cfg, sel)
    fg = make_graph(cfg["shape"], seed=0, device=torch.device("cpu"))
    net = out[0].net
    net.eval()
    out[0].trace = None
    with torch.no_grad():
        logits = net(O.EdgeList(fg.src, fg.dst(), fg.n_nodes, fg.n_nodes), fg.feat)
    want = ranks[0]["eval_logits"]
    assert logits.shape == want.shape == (fg.n_nodes, fg.n_class)
    assert _rel(logits, want) < 1e-5


def test_oracle_active_set_override():
    """``OracleRank.epoch(relu_masks=...)`` (the kink-aware leg of tests/harness.py): prescribing the oracle's OWN
    active sets changes nothing; flipping the entry closest to the kink is counted, its distance from zero reported,
    and moves the gradients -- which is exactly the disagreement a forward 1e-6 away can produce."""
    import bns_gcn_b200  # noqa: F401
    from bns_gcn_b200.data import make_graph, partition_graph
    from oracle import bns_oracle as O
    fg = make_graph("tiny", seed=0, device=torch.device("cpu"))
    parts = partition_graph(fg, 2, "random", seed=0, device=torch.device("cpu"))
    E = 2

    def run(masks_per_epoch):
        def fn(comm, r):
            rk = O.OracleRank(O.RankInput.from_partition(parts[r]), comm, model="graphsage", n_layers=3, n_hidden=16,
                              sampling_rate=0.5, dropout=0.0, seed=0)
            zs = []
            for e in range(E):
                rm = None if masks_per_epoch is None else masks_per_epoch[e][r]
                rk.epoch(rng=np.random.RandomState(100 * e + r), trace=True, relu_masks=rm)
                zs.append({i: rk.trace[f"z{i}"] for i in range(2)})
            return zs, [p.grad.clone() for p in rk.net.parameters()], dict(rk.kink)
        return O.run_threads(2, fn)

    base = run(None)
    own = [[{i: base[r][0][e][i] > 0 for i in range(2)} for r in range(2)] for e in range(E)]
    same = run(own)
    assert all(s[2]["flips"] == 0 for s in same)
    for a, b in zip(base[0][1], same[0][1]):
        assert _rel(a, b) < 1e-6
    # flip the entry of epoch 0 / rank 0 / norm 0 that is closest to zero
    z = base[0][0][0][0]
    k = int(z.abs().argmin())
    flipped = [[{i: m.clone() for i, m in own[e][r].items()} for r in range(2)] for e in range(E)]
    flipped[0][0][0].view(-1)[k] = ~flipped[0][0][0].view(-1)[k]
    # later epochs: keep the oracle's own (then current) sign except that the prescription must stay self-consistent,
    # so only epoch 0 is prescribed
    flipped = [flipped[0]] + [[None, None] for _ in range(E - 1)]
    moved = run(flipped)
    assert moved[0][2]["flips"] == 1 and moved[1][2]["flips"] == 0
    assert abs(moved[0][2]["max_abs_z"] - float(z.abs().min())) < 1e-12
    assert max(_rel(a, b) for a, b in zip(moved[0][1], base[0][1])) > 1e-7          # the kink matters


def test_harness_kink_retry_logic_with_a_stub_product(monkeypatch):
    """tests/harness.run_parity_case's kink-aware leg, exercised on the CPU: the CUDA run is replaced by a stub that
    computes the oracle's math but takes the OTHER side of the kink at the pre-activation closest to zero (what a
    forward that differs in the last bits does).  The first comparison must fail, the retry on the stub's active sets
    must bring the error back under the bar, and the report must say one entry was switched, at that distance."""
    import bns_gcn_b200  # noqa: F401
    from oracle import bns_oracle as O, philox
    from tests import harness as H

    state = {}

    def stub_run_product(parts, args, device, n_epochs, selected_per_epoch=None, capture=True, capture_masks=False):
        P = len(parts)

        def once(masks_per_epoch):
            def fn(comm, r):
                rk = O.OracleRank(O.RankInput.from_partition(parts[r]), comm, model=args.model, n_layers=args.n_layers,
                                  n_hidden=args.n_hidden, sampling_rate=args.sampling_rate, dropout=0.0, seed=args.seed)
                peers = [j for j in range(P) if j != r]
                losses, sel_log, hop_log, z_log = [], [], [], []
                for e in range(n_epochs):
                    drawn = philox.sample_boundary([rk.boundary[j].numpy() for j in peers],
                                                   [rk.send_size[j] for j in peers], args.sampler_seed, e)
                    sel = [None] * P
                    for i, j in enumerate(peers):
                        sel[j] = torch.from_numpy(np.ascontiguousarray(drawn[i])).long()
                    rm = None if masks_per_epoch is None else masks_per_epoch[e][r]
                    losses.append(rk.epoch(selected=sel, trace=True, relu_masks=rm))
                    sel_log.append(sel)
                    hop_log.append(list(rk.one_hops))
                    z_log.append({i: rk.trace[f"z{i}"] for i in range(args.n_layers - 1)})
                return {"loss": losses, "selected": sel_log, "one_hops": hop_log, "z": z_log,
                        "layers": {k: v for k, v in rk.trace.items() if k.startswith("layer")},
                        "logits": rk.trace["logits"], "grads": [q.grad.detach().clone() for q in rk.net.parameters()],
                        "params": [q.detach().clone() for q in rk.net.parameters()], "boundary": rk.boundary,
                        "send_size": rk.send_size, "feat0": rk.feat}
            return O.run_threads(P, fn)

        plain = once(None)
        z = plain[0]["z"][0][0]
        k = int(z.abs().argmin())
        state["abs_z"] = float(z.abs().min())
        first = [{i: (plain[r]["z"][0][i] > 0) for i in plain[r]["z"][0]} for r in range(P)]
        first[0][0] = first[0][0].clone()
        first[0][0].view(-1)[k] = ~first[0][0].view(-1)[k]
        out = once([first] + [[None] * P for _ in range(n_epochs - 1)])
        for r in range(P):       # the active sets this "implementation" actually took
            out[r]["relu_masks"] = [first[r]] + [{i: out[r]["z"][e][i] > 0 for i in out[r]["z"][e]}
                                                 for e in range(1, n_epochs)]
        state["calls"] = state.get("calls", 0) + 1
        return out

    monkeypatch.setattr(H, "run_product", stub_run_product)
    monkeypatch.setattr(H, "KINK_MARGIN", 1e-3)      # the 600-node graph's closest entry sits at ~1e-4, not ~1e-6
    res = H.run_parity_case(shape="tiny", n_parts=2, model="graphsage", sampling_rate=0.5, n_epochs=2, device="cpu")
    assert state["calls"] == 2, "the mismatch must trigger the second, mask-recording run"
    assert res["kink"] is not None and res["kink"]["flips"] == 1
    assert abs(res["kink"]["max_abs_z"] - state["abs_z"]) < 1e-9 and res["kink"]["max_abs_z"] < H.KINK_MARGIN
    assert res["kink"]["max_rel_err_before"] >= H.KINK_TRIGGER > res["max_rel_err"]
    assert res["index_sets_equal"]
    # a sign disagreement FAR from zero is a real forward error: with the margin below this entry's distance the retry
    # must not paper over it
    monkeypatch.setattr(H, "KINK_MARGIN", 1e-5)
    res = H.run_parity_case(shape="tiny", n_parts=2, model="graphsage", sampling_rate=0.5, n_epochs=2, device="cpu")
    assert res["kink"]["flips"] == 1 and res["max_rel_err"] >= H.KINK_TRIGGER


def test_c_spmm_property_random_shapes_vs_dense():
    """SURVEY §4 pin 4 as a property test (hypothesis): the oracle's C SpMM (oracle/spmm_ref.c, DGL's
    ``update_all(copy_u, sum)``) and its transpose against a dense 0/1-count matrix product, over random shapes
    including empty rows, empty graphs, duplicate edges and width-1 features."""
    from hypothesis import given, settings, strategies as st
    from oracle import bns_oracle as O

    @settings(max_examples=40, deadline=None)
    @given(n_src=st.integers(1, 40), n_dst=st.integers(1, 30), n_edges=st.integers(0, 300), width=st.integers(1, 9),
           seed=st.integers(0, 10 ** 6))
    def prop(n_src, n_dst, n_edges, width, seed):
        g = torch.Generator().manual_seed(seed)
        u = torch.randint(0, n_src, (n_edges,), generator=g)
        v = torch.randint(0, n_dst, (n_edges,), generator=g)
        x = torch.randn(n_src, width, generator=g, dtype=torch.float64).float().requires_grad_()
        a = torch.zeros(n_dst, n_src, dtype=torch.float64)
        a.index_put_((v, u), torch.ones(n_edges, dtype=torch.float64), accumulate=True)       # multi-edges count
        y = O.CopyUSum.apply(O.EdgeList(u, v, n_src, n_dst), x)
        want = a @ x.detach().double()
        assert y.shape == (n_dst, width)
        assert (y.detach().double() - want).abs().max() <= 1e-5 * (1 + want.abs().max())
        w = torch.randn(n_dst, width, generator=g)
        (y * w).sum().backward()
        want_g = a.t() @ w.double()
        assert (x.grad.double() - want_g).abs().max() <= 1e-5 * (1 + want_g.abs().max())

    prop()
