"""GPU parity of every C-ABI compute entry point against the oracle (oracle/spmm_ref.c, oracle/philox.py,
torch CPU fp32), on seeded inputs the oracle finishes in seconds.  Tolerances: index / integer work is
bit-exact; f32 sums 1e-5 relative (north_star allows 1e-4 on layer outputs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _rand_csr(n_rows, n_cols, avg_deg, seed, heavy=0, empty_frac=0.1):
    g = torch.Generator().manual_seed(seed)
    deg = torch.poisson(torch.full((n_rows,), float(avg_deg)), generator=g).long()
    deg[torch.rand(n_rows, generator=g) < empty_frac] = 0
    for i in range(heavy):                       # rows far longer than one chunk
        deg[(i * 7919) % n_rows] = 3000 + 1700 * i
    deg.clamp_(max=n_cols * 4)
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64)
    indptr[1:] = deg.cumsum(0)
    nnz = int(indptr[-1])
    idx = torch.randint(0, n_cols, (nnz,), generator=g, dtype=torch.int64)
    return indptr, idx


def _ref_spmm(indptr, idx, x, row_scale=None, col_scale=None, col_map=None, n_direct=None, row_map=None,
              n_out=None, y0=None):
    """Oracle: C SpMM on the (mapped, filtered) edge list, scalings applied the way the reference does."""
    from oracle import bns_oracle as O
    n_rows = indptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n_rows), indptr[1:] - indptr[:-1])
    cols = idx.clone()
    xrow = cols.clone()
    if col_map is not None:
        m = cols >= n_direct
        xrow[m] = col_map[cols[m] - n_direct].long()
    keep = xrow >= 0
    orow = rows.clone()
    if row_map is not None:
        orow = row_map[rows].long()
        keep &= orow >= 0
    n_out = n_out if n_out is not None else n_rows
    # scale sources per EDGE (col_scale is indexed by the original column id)
    if col_scale is not None:
        contrib = x[xrow[keep]] * col_scale[cols[keep]].unsqueeze(1)
        out = torch.zeros(n_out, x.shape[1]).index_add_(0, orow[keep], contrib)
    else:
        g = O.EdgeList(xrow[keep], orow[keep], x.shape[0], n_out)
        out = O.CopyUSum.apply(g, x)
    if row_scale is not None:
        rs = torch.zeros(n_out)
        if row_map is None:
            rs = row_scale
        else:
            ok = row_map >= 0
            rs[row_map[ok].long()] = row_scale[ok]
        out = out * rs.unsqueeze(1)
    if y0 is not None:
        out = out + y0
    return out


def _relerr(a, b):
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


@pytest.mark.parametrize("F", [256, 128, 64, 4, 100, 602, 41, 1, 300, 44, 48, 40])
def test_spmm_plain(built, F):
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    indptr, idx = _rand_csr(700, 900, 12, seed=F, heavy=2)
    x = torch.randn(900, F, generator=torch.Generator().manual_seed(F + 1))
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), 900)
    assert g.n_split_rows >= 2
    y = ops.spmm(g, x.to(dev)).cpu()
    ref = _ref_spmm(indptr, idx, x)
    assert _relerr(y, ref) < RTOL
    # rows without entries must be written as zeros
    empty = (indptr[1:] - indptr[:-1]) == 0
    assert empty.any() and torch.all(y[empty] == 0)


@pytest.mark.parametrize("F", [256, 604, 128, 44])
@pytest.mark.parametrize("slab", [0, 256, 128, 64, 32])
def test_spmm_every_slab_variant(built, F, slab):
    """Every column-slab instantiation of spmm_kernel against oracle/spmm_ref.c -- including the ones the default
    heuristic only picks on graphs too large for a unit test: <4,32,1> with n_tiles > 1 (F = 256 as two 128-float slabs,
    the instantiation behind the headline bench number), <4,16,1> and <4,8,1> multi-tile (sub-warp row groups), with the
    GUARD path where the slab does not divide F (604, 44).  Rows longer than one chunk exercise the partial sums."""
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    indptr, idx = _rand_csr(900, 1100, 14, seed=F + slab, heavy=3)
    x = torch.randn(1100, F, generator=torch.Generator().manual_seed(F + slab + 1))
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), 1100)
    assert g.n_split_rows >= 3
    y = ops.spmm(g, x.to(dev), slab=slab).cpu()
    ref = _ref_spmm(indptr, idx, x)
    assert _relerr(y, ref) < RTOL
    # full-warp slabs (256 / 128) keep the per-row summation order of the unblocked kernel: bit-identical; sub-warp
    # row groups (64 / 32) sum the entries of a chunk in a different order: equal to rounding
    y_full = ops.spmm(g, x.to(dev), slab=256).cpu()
    if slab in (128, 256):
        assert torch.equal(y, y_full)
    else:
        assert _relerr(y, y_full) < 1e-6


@pytest.mark.parametrize("slab", [256, 128, 64])
@pytest.mark.parametrize("F", [256, 604, 128])
def test_spmm_slab_variants_with_maps_scales_accumulate(built, F, slab):
    """The per-epoch forms (slot-mapped halo columns + col/row scales + accumulate; row-mapped backward) under a forced
    slab: the MAP / CSCALE instantiations of the blocked kernels."""
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(31 + F + slab)
    n_rows, n_direct, n_halo, n_slab = 500, 500, 420, 150
    indptr, idx = _rand_csr(n_rows, n_direct + n_halo, 18, seed=F + slab, heavy=2)
    slot = torch.full((n_halo,), -1, dtype=torch.int32)
    chosen = torch.randperm(n_halo, generator=gen)[:n_slab]
    slot[chosen] = torch.randperm(n_slab, generator=gen).int()
    x = torch.randn(n_direct + n_slab, F, generator=gen)
    cs = torch.rand(n_direct + n_halo, generator=gen) + 0.5
    rs = torch.rand(n_rows, generator=gen) + 0.5
    col_map = torch.where(slot >= 0, slot + n_direct, slot)
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), n_direct + n_halo)
    y0 = torch.randn(n_rows, F, generator=gen)
    y = y0.clone().to(dev)
    ops.spmm(g, x.to(dev), y, row_scale=rs.to(dev), col_scale=cs.to(dev), col_map=col_map.to(dev),
             n_direct=n_direct, accumulate=True, slab=slab)
    ref = _ref_spmm(indptr, idx, x, row_scale=rs, col_scale=cs, col_map=col_map, n_direct=n_direct) + y0
    assert _relerr(y.cpu(), ref) < RTOL
    ind2, idx2 = _rand_csr(n_halo, n_rows, 20, seed=F + slab + 100, heavy=1, empty_frac=0.0)
    dy = torch.randn(n_rows, F, generator=gen)
    g2 = ops.DeviceGraph.from_csr(ind2.to(dev), idx2.int().to(dev), n_rows)
    out = torch.full((n_slab, F), 7.0, device=dev)
    ops.spmm(g2, dy.to(dev), out, row_map=slot.to(dev), row_scale=cs[n_direct:].to(dev), slab=slab)
    ref2 = _ref_spmm(ind2, idx2, dy, row_scale=cs[n_direct:], row_map=slot, n_out=n_slab)
    assert _relerr(out.cpu(), ref2) < RTOL


def test_spmm_roundtrip_csr_and_transpose(built):
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    indptr, idx = _rand_csr(500, 300, 9, seed=3, heavy=1)
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), 300)
    ip, ix = g.csr()
    assert torch.equal(ip.cpu(), indptr) and torch.equal(ix.cpu().long(), idx)        # bit-exact copy
    gt = g.transpose()
    assert (gt.n_rows, gt.n_cols, gt.nnz) == (300, 500, g.nnz)
    tp, tx = (t.cpu() for t in gt.csr())
    rows = torch.repeat_interleave(torch.arange(500), indptr[1:] - indptr[:-1])
    order = torch.argsort(idx * 500 + rows, stable=True)                                # by column, rows ascending
    assert torch.equal(tx.long(), rows[order])
    cnt = torch.bincount(idx, minlength=300)
    assert torch.equal(tp[1:] - tp[:-1], cnt)
    # backward of the aggregation == SpMM on the transpose == oracle's reversed-graph SpMM
    dy = torch.randn(500, 64, generator=torch.Generator().manual_seed(5))
    dx = ops.spmm(gt, dy.to(dev)).cpu()
    from oracle import bns_oracle as O
    e = O.EdgeList(idx, rows, 300, 500)
    tip, tcols = e.csr_t()
    ref = O._spmm(tip, tcols, dy, 300)
    assert _relerr(dx, ref) < RTOL


@pytest.mark.parametrize("F", [256, 128, 36, 7, 44])
def test_spmm_scales_maps_accumulate(built, F):
    """The per-epoch form: halo columns resolved through a slot map (-1 = unsampled), GCN-style col / row
    scales, accumulation on top of the inner-edge pass, and the row-mapped backward."""
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(11 + F)
    n_rows, n_direct, n_halo, n_slab = 400, 400, 350, 120
    indptr, idx = _rand_csr(n_rows, n_direct + n_halo, 15, seed=F, heavy=1)
    slot = torch.full((n_halo,), -1, dtype=torch.int32)
    chosen = torch.randperm(n_halo, generator=gen)[:n_slab]
    slot[chosen] = torch.randperm(n_slab, generator=gen).int()
    x = torch.randn(n_direct + n_slab, F, generator=gen)
    cs = torch.rand(n_direct + n_halo, generator=gen) + 0.5
    rs = torch.rand(n_rows, generator=gen) + 0.5
    col_map = torch.where(slot >= 0, slot + n_direct, slot)        # absolute row of x
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), n_direct + n_halo)
    y0 = torch.randn(n_rows, F, generator=gen)
    y = y0.clone().to(dev)
    ops.spmm(g, x.to(dev), y, row_scale=rs.to(dev), col_scale=cs.to(dev), col_map=col_map.to(dev),
             n_direct=n_direct, accumulate=True)
    ref = _ref_spmm(indptr, idx, x, row_scale=rs, col_scale=cs, col_map=col_map, n_direct=n_direct) + y0
    assert _relerr(y.cpu(), ref) < RTOL
    # row-mapped (backward over sampled halo rows): rows = halo nodes, output row = slab slot
    ind2, idx2 = _rand_csr(n_halo, n_rows, 20, seed=F + 100, heavy=1, empty_frac=0.0)
    dy = torch.randn(n_rows, F, generator=gen)
    g2 = ops.DeviceGraph.from_csr(ind2.to(dev), idx2.int().to(dev), n_rows)
    out = torch.full((n_slab, F), 7.0, device=dev)
    ops.spmm(g2, dy.to(dev), out, row_map=slot.to(dev), row_scale=cs[n_direct:].to(dev), col_scale=rs.to(dev))
    ref2 = _ref_spmm(ind2, idx2, dy, row_scale=cs[n_direct:], col_scale=rs, row_map=slot, n_out=n_slab)
    assert _relerr(out.cpu(), ref2) < RTOL


def test_spmm_edge_cases(built):
    from bns_gcn_b200 import ops, _lib
    dev = torch.device("cuda:0")
    # empty graph (no rows), graph with rows but no entries
    g0 = ops.DeviceGraph.from_csr(torch.zeros(1, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int32, device=dev), 5)
    assert ops.spmm(g0, torch.randn(5, 8, device=dev)).shape == (0, 8)
    g1 = ops.DeviceGraph.from_csr(torch.zeros(4, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int32, device=dev), 5)
    assert torch.all(ops.spmm(g1, torch.randn(5, 8, device=dev)) == 0)
    # non-contiguous leading dimension (a column slice) takes the scalar / strided path
    indptr, idx = _rand_csr(64, 64, 5, seed=1)
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), 64)
    big = torch.randn(64, 40)
    y = ops.spmm(g, big.to(dev)[:, 3:35]).cpu()
    assert _relerr(y, _ref_spmm(indptr, idx, big[:, 3:35].contiguous())) < RTOL
    # out-of-range column index is rejected at creation, with a message
    bad = idx.clone()
    bad[0] = 64
    with pytest.raises(_lib.BnsError, match="outside"):
        ops.DeviceGraph.from_csr(indptr.to(dev), bad.int().to(dev), 64)
    with pytest.raises(_lib.BnsError):
        ops.spmm(g, torch.randn(64, 8))          # CPU tensor: no CPU path


def test_aggregate_autograd_matches_oracle(built):
    from bns_gcn_b200 import ops
    from oracle import bns_oracle as O
    dev = torch.device("cuda:0")
    indptr, idx = _rand_csr(300, 300, 10, seed=9, heavy=1, empty_frac=0.0)
    rows = torch.repeat_interleave(torch.arange(300), indptr[1:] - indptr[:-1])
    deg = (indptr[1:] - indptr[:-1]).clamp(min=1).float()
    x = torch.randn(300, 128, generator=torch.Generator().manual_seed(2), requires_grad=True)
    w = torch.randn(300, 128, generator=torch.Generator().manual_seed(3))
    ref = O.CopyUSum.apply(O.EdgeList(idx, rows, 300, 300), x) / deg.unsqueeze(1)
    (ref * w).sum().backward()
    xg = x.detach().to(dev).requires_grad_(True)
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), 300)
    y = ops.AggregateSum.apply(xg, g, (1.0 / deg).to(dev), None)
    (y * w.to(dev)).sum().backward()
    assert _relerr(y.detach().cpu(), ref.detach()) < RTOL
    assert _relerr(xg.grad.cpu(), x.grad) < RTOL


@pytest.mark.parametrize("F", [256, 41, 602])
def test_gather_scatter(built, F):
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(F)
    h = torch.randn(500, F, generator=gen)
    idx = torch.randperm(500, generator=gen)[:123]
    ratio = 123 / 500
    out = ops.gather_div(h.to(dev), idx.to(dev), ratio).cpu()
    assert torch.equal(out, h[idx] / ratio)                       # true division: bit-exact vs torch CPU
    grad = torch.randn(500, F, generator=gen)
    src = torch.randn(123, F, generator=gen)
    ref = grad.clone()
    ref[idx] += src / ratio
    got = ops.scatter_add_div(grad.clone().to(dev), idx.to(dev), src.to(dev), ratio).cpu()
    assert torch.equal(got, ref)
    # empty selection is a no-op
    e = torch.empty(0, dtype=torch.int64, device=dev)
    assert ops.gather_div(h.to(dev), e, 1.0).shape == (0, F)


def test_sampler_exact_vs_philox_replay(built):
    from bns_gcn_b200 import ops
    from oracle import philox
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    boundary = [None, torch.sort(torch.randperm(5000, generator=gen)[:1733])[0], torch.empty(0, dtype=torch.int64),
                torch.sort(torch.randperm(9000, generator=gen)[:4000])[0]]
    send = [0, int(0.1 * 1733), 0, int(0.37 * 4000)]
    s = ops.BoundarySampler(boundary, send, dev)
    for seed, off in [(0, 0), (12345, 7), (2**40 + 3, 2**33 + 1)]:
        _, views = s.sample(seed, off)
        ref = philox.sample_boundary([b.numpy() for b in boundary if b is not None], [send[j] for j in s.peers], seed, off)
        for i, j in enumerate(s.peers):
            assert torch.equal(views[j].cpu(), torch.from_numpy(ref[i])), (seed, off, j)      # exact, ordered
            v = views[j].cpu()
            assert v.numel() == send[j] and v.unique().numel() == v.numel()                    # no duplicates
            assert torch.isin(v, boundary[j]).all()                                            # subset
    # different epochs give different samples; same (seed, offset) is reproducible
    a = s.sample(1, 5)[0].clone()
    assert torch.equal(a, s.sample(1, 5)[0]) and not torch.equal(a, s.sample(1, 6)[0])


def test_sampler_uniform_inclusion(built):
    """chi-square on inclusion counts: every boundary node is sampled with probability k/b."""
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    b, k, trials = 200, 50, 2000
    s = ops.BoundarySampler([None, torch.arange(b)], [0, k], dev)
    cnt = torch.zeros(b)
    for t in range(trials):
        cnt += torch.bincount(s.sample(99, t)[0].cpu(), minlength=b)
    exp = trials * k / b
    chi2 = (((cnt - exp) ** 2) / (exp * (1 - k / b))).sum().item()      # ~ chi2(b-1): mean 199, sd ~20
    assert 120 < chi2 < 290, chi2


def test_halo_slot_update(built):
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    n_in, n_halo, part = 100, 60, 80
    gen = torch.Generator().manual_seed(4)
    pos = torch.full((part,), -1, dtype=torch.int64)
    owned = torch.randperm(part, generator=gen)[:n_halo]
    pos[owned] = n_in + torch.arange(n_halo)
    one_hops = owned[torch.randperm(n_halo, generator=gen)[:25]]
    slot = torch.empty(n_halo, dtype=torch.int32, device=dev)
    ops.fill_i32(slot, -1)
    ops.halo_slot_update(pos.to(dev), one_hops.to(dev), n_in, 1000, slot)
    ref = torch.full((n_halo,), -1, dtype=torch.int32)
    ref[pos[one_hops] - n_in] = 1000 + torch.arange(25, dtype=torch.int32)
    assert torch.equal(slot.cpu(), ref)


def test_dense_3xtf32_is_fp32_accurate(built):
    """The error-compensated tensor-core linear stays at f32-level accuracy (vs an f64 reference), forward and
    backward; a single TF32 pass would be ~1e-3."""
    from bns_gcn_b200.module import dense
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, 1204, generator=g).to(dev).requires_grad_(True)
    w = (torch.rand(256, 1204, generator=g) - 0.5).to(dev).requires_grad_(True)
    b = torch.randn(256, generator=g).to(dev).requires_grad_(True)
    y = dense._Linear3x.apply(x, w, b)
    dy = torch.randn(4096, 256, generator=g).to(dev)
    y.backward(dy)
    xd, wd, bd = x.detach().double(), w.detach().double(), b.detach().double()
    ref = xd @ wd.t() + bd
    assert _relerr(y.detach().double().cpu(), ref.cpu()) < 1e-5
    assert _relerr(x.grad.double().cpu(), (dy.double() @ wd).cpu()) < 1e-5
    assert _relerr(w.grad.double().cpu(), (dy.double().t() @ xd).cpu()) < 1e-5
    assert _relerr(b.grad.double().cpu(), dy.double().sum(0).cpu()) < 1e-5
    # plain fp32 cuBLAS for comparison: same order of magnitude of error
    y32 = torch.nn.functional.linear(x.detach(), w.detach(), b.detach())
    assert _relerr(y.detach().cpu(), y32.cpu()) < 1e-5


@pytest.mark.parametrize("F,p", [(256, 0.0), (256, 0.5), (64, 0.3), (600, 0.5), (16, 0.0)])
def test_fused_layernorm_relu_dropout(built, F, p):
    """ops.LnReluDropout == dropout(relu(layer_norm(x))) forward and backward (mask recovered from the output),
    mask keep-rate ~ 1-p, masks differ across offsets and repeat for the same (seed, offset)."""
    import torch.nn.functional as Fn
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(F)
    n = 3000
    x = (torch.randn(n, F, generator=g) * 2 + 0.3).to(dev).requires_grad_(True)
    gamma = (torch.rand(F, generator=g) + 0.5).to(dev).requires_grad_(True)
    beta = (torch.randn(F, generator=g) * 0.2).to(dev).requires_grad_(True)
    dy = torch.randn(n, F, generator=g).to(dev)
    ops.RNG.update(seed=123, offset=5, offset_dev=None)
    y = ops.LnReluDropout.apply(x, gamma, beta, 1e-5, p, 77)
    y.backward(dy)
    got = (y.detach().clone(), x.grad.clone(), gamma.grad.clone(), beta.grad.clone())
    # reference with the same mask
    xr, gr, br = (t.detach().clone().requires_grad_(True) for t in (x, gamma, beta))
    z = Fn.relu(Fn.layer_norm(xr, (F,), gr, br, 1e-5))
    if p > 0:
        mask = ((y.detach() != 0) | (z.detach() <= 0)).float()
        keep = mask[z.detach() > 0].mean().item()
        assert abs(keep - (1 - p)) < 0.01, keep
        ref = z * mask / (1 - p)
    else:
        ref = z
    ref.backward(dy)
    for a, b in zip(got, (ref.detach(), xr.grad, gr.grad, br.grad)):
        assert _relerr(a.cpu(), b.cpu()) < 2e-5
    if p > 0:
        y2 = ops.LnReluDropout.apply(x.detach(), gamma.detach(), beta.detach(), 1e-5, p, 77)
        assert torch.equal(y2, y.detach())                            # same (seed, offset) -> same mask
        off_dev = torch.tensor([1], dtype=torch.int64, device=dev)    # 5 + 1: offset read from the device
        ops.RNG.update(offset=5, offset_dev=off_dev)
        y3 = ops.LnReluDropout.apply(x.detach(), gamma.detach(), beta.detach(), 1e-5, p, 77)
        ops.RNG.update(offset=6, offset_dev=None)
        y4 = ops.LnReluDropout.apply(x.detach(), gamma.detach(), beta.detach(), 1e-5, p, 77)
        assert torch.equal(y3, y4) and not torch.equal(y3, y.detach())
    ops.RNG.update(seed=0, offset=0, offset_dev=None)


@pytest.mark.parametrize("F", [64, 256, 44])
def test_weighted_spmm_perm_and_sddmm(built, F):
    """GAT primitives: per-entry weights in the SpMM, the transpose's entry permutation, and the SDDMM dot."""
    from bns_gcn_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(F)
    n_rows, n_direct, n_halo, n_slab = 300, 300, 200, 70
    indptr, idx = _rand_csr(n_rows, n_direct + n_halo, 14, seed=F + 3, heavy=1)
    nnz = idx.numel()
    rows = torch.repeat_interleave(torch.arange(n_rows), indptr[1:] - indptr[:-1])
    w = torch.rand(nnz, generator=gen)
    slot = torch.full((n_halo,), -1, dtype=torch.int32)
    chosen = torch.randperm(n_halo, generator=gen)[:n_slab]
    slot[chosen] = torch.randperm(n_slab, generator=gen).int()
    col_map = torch.where(slot >= 0, slot + n_direct, slot)
    x = torch.randn(n_direct + n_slab, F, generator=gen)
    g = ops.DeviceGraph.from_csr(indptr.to(dev), idx.int().to(dev), n_direct + n_halo, 64)
    y = ops.spmm(g, x.to(dev), edge_weight=w.to(dev), col_map=col_map.to(dev), n_direct=n_direct).cpu()
    xrow = idx.clone()
    m = idx >= n_direct
    xrow[m] = col_map[idx[m] - n_direct].long()
    keep = xrow >= 0
    ref = torch.zeros(n_rows, F).index_add_(0, rows[keep], x[xrow[keep]] * w[keep].unsqueeze(1))
    assert _relerr(y, ref) < RTOL
    # transpose permutation: entry k of g^T is entry perm[k] of g
    gt = g.transpose()
    perm = gt.perm().cpu().long()
    tp, tx = (t.cpu() for t in gt.csr())
    assert torch.equal(tx.long(), rows[perm])
    tcols = torch.repeat_interleave(torch.arange(gt.n_rows), tp[1:] - tp[:-1])
    assert torch.equal(tcols, idx[perm])
    # weighted transpose SpMM == autograd of the forward
    dy = torch.randn(n_rows, F, generator=gen)
    dx = ops.spmm(gt, dy.to(dev), edge_weight=w[perm].to(dev)).cpu()
    ref_dx = torch.zeros(n_direct + n_halo, F).index_add_(0, idx, dy[rows] * w.unsqueeze(1))
    assert _relerr(dx, ref_dx) < RTOL
    # SDDMM dot with the column map: d w_k = <dy[row_k], x[xrow_k]>, 0 for skipped entries
    if F % 4 == 0:
        dw = ops.sddmm_dot(g, dy.to(dev), x.to(dev), col_map=col_map.to(dev), n_direct=n_direct).cpu()
        ref_dw = torch.zeros(nnz)
        ref_dw[keep] = (dy[rows[keep]] * x[xrow[keep]]).sum(1)
        assert _relerr(dw, ref_dw) < RTOL
        assert torch.all(dw[~keep] == 0)


def test_dense_bf16x3_is_fp32_accurate(built):
    """Three-way bf16 split + six tensor-core GEMMs with f32 accumulation: f32-level accuracy vs an f64 reference."""
    from bns_gcn_b200.module import dense
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4096, 1204, generator=g).to(dev).requires_grad_(True)
    w = (torch.rand(256, 1204, generator=g) - 0.5).to(dev).requires_grad_(True)
    b = torch.randn(256, generator=g).to(dev).requires_grad_(True)
    x3 = dense._split3(x.detach())
    assert torch.equal((x3[0].float() + x3[1].float()) + x3[2].float(), x.detach()) or \
        _relerr(((x3[0].float() + x3[1].float()) + x3[2].float()).cpu(), x.detach().cpu()) < 1e-7
    y = dense._LinearBf16x3.apply(x, w, b)
    dy = torch.randn(4096, 256, generator=g).to(dev)
    y.backward(dy)
    xd, wd, bd = x.detach().double(), w.detach().double(), b.detach().double()
    assert _relerr(y.detach().double().cpu(), (xd @ wd.t() + bd).cpu()) < 2e-6
    assert _relerr(x.grad.double().cpu(), (dy.double() @ wd).cpu()) < 2e-6
    assert _relerr(w.grad.double().cpu(), (dy.double().t() @ xd).cpu()) < 2e-6
