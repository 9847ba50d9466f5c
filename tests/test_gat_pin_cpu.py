"""Pins for ``oracle.GATConvRef`` (the restatement of ``dgl.nn.GATConv``, which module/model.py:102 of the reference
constructs and which is not vendored): hand-computed outputs on a 3-source / 2-destination bipartite graph, derived on
paper from the layer's published definition (DGL 0.9 python/dgl/nn/pytorch/conv/gatconv.py; Velickovic et al. 2018)

    ft = W h        el_u = <ft_u, attn_l>        er_v = <ft_v, attn_r>
    e_uv = LeakyReLU_0.2(el_u + er_v)            a_uv = exp(e_uv) / sum_{u' -> v} exp(e_u'v)
    rst_v = sum_{u -> v} a_uv ft_u + bias

The numbers below are literals (multiples of ln 2 chosen so that every softmax is a ratio of small integers), not the
output of any implementation.  The CUDA path is compared with this oracle by tests/test_parity_gpu.py::test_gat_*, and
its fused attention kernels with an independent f64 restatement by tests/test_kernels_gpu.py::test_fused_gat_*."""
import math

import torch

LN2 = math.log(2.0)


def _layer(H, Fo, in_feats, W, attn_l, attn_r, bias=None):
    from oracle import bns_oracle as O
    conv = O.GATConvRef(in_feats, Fo, H, 0.0, 0.0)
    with torch.no_grad():
        conv.fc.weight.copy_(torch.tensor(W, dtype=torch.float32))
        conv.attn_l.copy_(torch.tensor(attn_l, dtype=torch.float32).view(1, H, Fo))
        conv.attn_r.copy_(torch.tensor(attn_r, dtype=torch.float32).view(1, H, Fo))
        conv.bias.copy_(torch.zeros(H * Fo) if bias is None else torch.tensor(bias, dtype=torch.float32))
    return conv.eval()


def _graph():
    from oracle import bns_oracle as O
    # edges u -> v:  0->0, 1->0, 0->1, 2->1 ; destination nodes are the first two source nodes (bipartite _U -> _V)
    return O.EdgeList(torch.tensor([0, 1, 0, 2]), torch.tensor([0, 0, 1, 1]), 3, 2)


def test_uniform_attention_is_the_mean_of_the_neighbours():
    """attn_l = attn_r = 0: every score is LeakyReLU(0) = 0, the softmax is uniform, rst is the neighbour mean."""
    conv = _layer(1, 2, 2, [[1, 0], [0, 1]], [0, 0], [0, 0], bias=[0.5, -1.0])
    h = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    out = conv(_graph(), (h, h[:2]))
    want = torch.tensor([[[2.0 + 0.5, 3.0 - 1.0]], [[3.0 + 0.5, 4.0 - 1.0]]])       # mean(rows 0,1), mean(rows 0,2), + bias
    assert torch.allclose(out, want, atol=1e-6), out


def test_source_scores_positive_and_negative_branch_of_the_leaky_relu():
    """attn_l = [1, 0], attn_r = 0, W = I: el_u = h_u[0].
    v = 0: sources 0 (el = 0 -> e = 0, exp = 1) and 1 (el = ln 2 -> e = ln 2, exp = 2): a = (1/3, 2/3)
           rst_0 = 1/3 [0, 1] + 2/3 [ln 2, 3] = [2 ln 2 / 3, 7/3]
    v = 1: sources 0 (exp = 1) and 2 (el = -5 ln 2 -> e = 0.2 * (-5 ln 2) = -ln 2, exp = 1/2): a = (2/3, 1/3)
           rst_1 = 2/3 [0, 1] + 1/3 [-5 ln 2, 9] = [-5 ln 2 / 3, 11/3]"""
    conv = _layer(1, 2, 2, [[1, 0], [0, 1]], [1, 0], [0, 0])
    h = torch.tensor([[0.0, 1.0], [LN2, 3.0], [-5 * LN2, 9.0]])
    out = conv(_graph(), (h, h[:2]))
    want = torch.tensor([[[2 * LN2 / 3, 7.0 / 3]], [[-5 * LN2 / 3, 11.0 / 3]]])
    assert torch.allclose(out, want, atol=1e-6), out


def test_two_heads_destination_scores_and_weight_matrix():
    """Two heads over a real fc: W (4 x 2) stacks head 0 = [[2, 0], [0, 1]] and head 1 = [[0, 1], [1, 0]] (a swap).
    head 0: attn_l = [1/2, 0], attn_r = 0 -> el_u = h_u[0] (= 2 h_u[0] / 2): the scores of the previous test.
            ft_u = [2 h_u0, h_u1]:  rst_0 = 1/3 [0, 1] + 2/3 [2 ln 2, 3] = [4 ln 2 / 3, 7/3]
                                    rst_1 = 2/3 [0, 1] + 1/3 [-10 ln 2, 9] = [-10 ln 2 / 3, 11/3]
    head 1: attn_l = 0, attn_r = [0, 7]: er_v is the same for all in-edges of v, so it cancels in the softmax: uniform.
            ft_u = [h_u1, h_u0]:    rst_0 = mean([1, 0], [3, ln 2]) = [2, ln 2 / 2]
                                    rst_1 = mean([1, 0], [9, -5 ln 2]) = [5, -5 ln 2 / 2]"""
    conv = _layer(2, 2, 2, [[2, 0], [0, 1], [0, 1], [1, 0]], [[0.5, 0], [0, 0]], [[0, 0], [0, 7]])
    h = torch.tensor([[0.0, 1.0], [LN2, 3.0], [-5 * LN2, 9.0]])
    out = conv(_graph(), (h, h[:2]))
    want = torch.tensor([[[4 * LN2 / 3, 7.0 / 3], [2.0, LN2 / 2]],
                         [[-10 * LN2 / 3, 11.0 / 3], [5.0, -5 * LN2 / 2]]])
    assert out.shape == (2, 2, 2)
    assert torch.allclose(out, want, atol=1e-6), out
