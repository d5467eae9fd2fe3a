"""CPU tests that pin the oracle (the reference ships no tests of its own -- SURVEY.md F3):
independent dense cross-check, closed-form known answers, fp64 gradcheck, literal-vs-collapsed
coefficient path, frozen golden vectors, and the PyG-1.7 utility semantics."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from helpers import random_batch_graph
from oracle import pyg17
from oracle.cheb import OracleChebConvDynamic, cheb_conv_dynamic, cheb_norm
from oracle.dense import dense_cheb, dense_coeff_scalar, dense_scaled_laplacian
from oracle.layers import OracleDiffTransformerEncoderLayer
from oracle.models import OracleEncoderGenGCN
import oracle.models as omodels

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@settings(max_examples=25, deadline=None)
@given(sizes=st.lists(st.integers(1, 9), min_size=1, max_size=5), K=st.integers(1, 4), seed=st.integers(0, 1000),
       float_batch=st.booleans())
def test_sparse_oracle_matches_dense_definition(sizes, K, seed, float_batch):
    ei, batch, R = random_batch_graph(seed, sizes, directed_extra=1)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, 3, generator=g, dtype=torch.float64)
    th = torch.randn(K, len(sizes), 3, 5, generator=g, dtype=torch.float64)
    bias = torch.randn(5, generator=g, dtype=torch.float64)
    b = batch.double() if float_batch else batch
    out = cheb_conv_dynamic(x, ei, th, batch=b, bias=bias).reshape(R, 5)
    ref = dense_cheb(x, ei, th, batch, bias)
    assert torch.allclose(out, ref, atol=1e-10)


def test_cheb_norm_double_self_loop_cancels():
    """A2: +1 from L and -1 from add_self_loops(fill=-1) leave a zero diagonal."""
    ei, batch, R = random_batch_graph(3, [6, 4])
    ei2, w = cheb_norm(ei, R, None, 'sym', torch.tensor(2.0), dtype=torch.float64)
    M = torch.zeros(R, R, dtype=torch.float64).index_put_((ei2[1], ei2[0]), w, accumulate=True)
    assert torch.allclose(M, dense_scaled_laplacian(ei, R), atol=1e-12)
    assert torch.allclose(torch.diagonal(M), torch.zeros(R, dtype=torch.float64))
    assert ei2.shape[1] == int((ei[0] != ei[1]).sum()) + 2 * R          # E' + two loops per node


def test_known_answers_isolated_and_p2():
    F, K = 3, 4
    x = torch.randn(5, F, dtype=torch.float64)
    batch = torch.tensor([0, 0, 1, 1, 1])
    ei = torch.tensor([[0, 1], [1, 0]])
    for k in range(K):
        th = torch.zeros(K, 2, F, F, dtype=torch.float64)
        th[k] = torch.eye(F, dtype=torch.float64)
        out = cheb_conv_dynamic(x, ei, th, batch=batch)
        iso = [x[2:], 0 * x[2:], -x[2:], 0 * x[2:]][k]                 # (x, 0, -x, 0)
        assert torch.allclose(out[2:], iso, atol=1e-12)
        p2 = x[:2] if k % 2 == 0 else -x[:2].flip(0)
        assert torch.allclose(out[:2], p2, atol=1e-12)


def test_known_answer_complete_graph():
    """K_n: L_hat = -(J - I)/(n-1); a constant vector is an eigenvector with eigenvalue -1."""
    n, F = 6, 2
    src, dst = zip(*[(i, j) for i in range(n) for j in range(n) if i != j])
    ei = torch.tensor([src, dst])
    x = torch.ones(n, F, dtype=torch.float64)
    th = torch.zeros(4, 1, F, F, dtype=torch.float64)
    th[:, 0] = torch.eye(F, dtype=torch.float64)
    out = cheb_conv_dynamic(x, ei, th, batch=torch.zeros(n, dtype=torch.long))
    # T_k(-1) = (-1)^k  ->  sum_k = 1 - 1 + 1 - 1 = 0
    assert torch.allclose(out, torch.zeros_like(out), atol=1e-12)


def test_gradcheck_fp64():
    ei, batch, R = random_batch_graph(5, [4, 3])
    g = torch.Generator().manual_seed(5)
    x = torch.randn(R, 2, generator=g, dtype=torch.float64, requires_grad=True)
    th = torch.randn(3, 2, 2, 2, generator=g, dtype=torch.float64, requires_grad=True)
    bias = torch.randn(2, generator=g, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b, c: cheb_conv_dynamic(a, ei, b, batch=batch, bias=c), (x, th, bias))


def test_attention_layer_gradcheck_fp64_and_rows():
    torch.manual_seed(0)
    layer = OracleDiffTransformerEncoderLayer(8, 2, 16, 0.0).double()
    lens = torch.tensor([5, 3])
    mask = torch.arange(5)[None, :] >= lens[:, None]
    a = torch.rand(2, 5, 5, dtype=torch.float64)
    pe = (a + a.transpose(1, 2)) * ((~mask)[:, :, None] & (~mask)[:, None, :])
    deg = torch.rand(2, 5, dtype=torch.float64) * (~mask)
    src = torch.randn(5, 2, 8, dtype=torch.float64, requires_grad=True)
    out, attn, heads = layer(src, pe=pe, degree=deg, src_key_padding_mask=mask, need_heads=True)
    assert attn.shape == (2, 2, 5, 5) and heads.shape == (2, 5, 2, 4) and out.shape == (5, 2, 8)
    real = (~mask)[:, None, :].expand(-1, 2, -1)
    assert torch.allclose(attn.sum(-1)[real], torch.ones(int(real.sum()), dtype=torch.float64))
    assert float(attn[1, :, :, 3:].abs().max()) == 0.0                   # masked keys get exactly 0
    assert torch.autograd.gradcheck(
        lambda s: layer(s, pe=pe, degree=deg, src_key_padding_mask=mask, need_heads=True)[0], (src,), atol=1e-5)


@pytest.mark.parametrize("lens", [[4, 2, 5], [1, 1], [7]])
def test_collapsed_coefficients_equal_literal_all_pairs_gcn(lens):
    """The closed form used by the CUDA path == models.py:240-287 restated literally."""
    torch.manual_seed(1)
    d, H = 8, 2
    enc = OracleEncoderGenGCN(d, H, OracleDiffTransformerEncoderLayer(d, H, 16, 0.0), 1).double()
    B, nmax = len(lens), max(lens)
    mask = torch.arange(nmax)[None, :] >= torch.tensor(lens)[:, None]
    a = torch.rand(B, H, nmax, nmax, dtype=torch.float64)
    a = a * (torch.rand(B, H, nmax, nmax) > 0.3)                         # exact zeros, some on the diagonal
    a = a * ((~mask)[:, None, :, None] & (~mask)[:, None, None, :])
    lit = enc.get_filter_coefficients(a, None, None, None, mask)
    enc.collapsed_coeff = True
    col = enc.get_filter_coefficients(a, None, None, None, mask)
    assert torch.allclose(lit, col, atol=1e-10)
    # and the scalar itself: GCNConv(ones) = s * colsum(W) + b
    n = lens[0]
    x = torch.ones(n, 4, dtype=torch.float64)
    W, b = torch.randn(4, 4, dtype=torch.float64), torch.randn(4, dtype=torch.float64)
    ag = a[0, 0, :n, :n]
    nz = ag.reshape(-1) != 0
    grid = torch.tensor(np.mgrid[0:n, 0:n].reshape(2, -1))
    out = pyg17.gcn_conv(x, grid[:, nz], ag.reshape(-1)[nz], W, b)
    s = dense_coeff_scalar(ag)
    assert torch.allclose(out, s.view(-1, 1) * W.sum(0).view(1, -1) + b, atol=1e-10)


def test_pyg17_utilities():
    ei = torch.tensor([[0, 1, 1, 2, 2], [1, 0, 1, 2, 0]])
    w = torch.tensor([1., 2., 3., 4., 5.])
    e2, w2 = pyg17.remove_self_loops(ei, w)
    assert e2.tolist() == [[0, 1, 2], [1, 0, 0]] and w2.tolist() == [1., 2., 5.]
    e3, w3 = pyg17.add_remaining_self_loops(ei, w, 1.0, 4)              # keeps existing loop weights
    assert e3.tolist() == [[0, 1, 2, 0, 1, 2, 3], [1, 0, 0, 0, 1, 2, 3]]
    assert w3.tolist() == [1., 2., 5., 1., 3., 4., 1.]
    e4, w4 = pyg17.add_self_loops(ei, w, -1.0, 3)                        # unconditional
    assert e4.shape[1] == 8 and w4[-3:].tolist() == [-1., -1., -1.]
    x = torch.arange(6.).view(3, 2)
    assert pyg17.global_mean_pool(x, torch.tensor([0, 0, 1])).tolist() == [[1., 2.], [4., 5.]]
    assert pyg17.degree(torch.tensor([0, 0, 2]), 4).tolist() == [2., 0., 1., 0.]
    el, wl = pyg17.get_laplacian(torch.tensor([[0, 1], [1, 0]]), None, 'sym', torch.float32, 3)
    L = torch.zeros(3, 3).index_put_((el[0], el[1]), wl, accumulate=True)
    assert torch.allclose(L, torch.tensor([[1., -1., 0.], [-1., 1., 0.], [0., 0., 1.]]))


def test_head_tiling_quirk_F4():
    """models.py:176-186: edge_index is NOT tiled per head -> heads >= 1 see isolated nodes."""
    torch.manual_seed(2)
    d, H = 8, 2
    enc = OracleEncoderGenGCN(d, H, OracleDiffTransformerEncoderLayer(d, H, 16, 0.0), 1)
    enc.collapsed_coeff = True
    n = 4
    fi = torch.tensor([[0, i] for i in range(n)])
    heads = torch.randn(1, n, H, d // H)
    coeff = torch.randn(H, 4 * (d // H) ** 2)
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    out_heads = heads.permute([2, 0, 1, 3]).reshape(H, n, d // H)
    fia = fi.repeat(H, 1)
    fia[:, 0] += torch.arange(H).repeat_interleave(n)
    bat = torch.arange(H).repeat_interleave(n).float()
    y = enc.filter(coeff, out_heads, ei, fia, bat, enc.spectral_gnns)
    th = coeff.reshape(H, 4, d // H, d // H)
    x1 = out_heads[1]
    assert torch.allclose(y[n:], x1 @ (th[1, 0] - th[1, 2]), atol=1e-5)  # T = (x, 0, -x, 0)


@pytest.mark.parametrize("name", ["cheb_case", "attention_case", "model_case"])
def test_golden_vectors(name):
    """Frozen outputs (tests/golden/make_golden.py) -- the oracle must keep reproducing them."""
    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    if name == "cheb_case":
        F, K = g['F'], g['K']
        m = OracleChebConvDynamic(F, F, K)
        m.bias.data.copy_(g['bias'])
        x, c = g['x'].clone().requires_grad_(), g['coeff'].clone().requires_grad_()
        out = m(x, g['edge_index'], c.reshape(-1, K, F, F).permute(1, 0, 2, 3), batch=g['batch'].float())
        out.backward(g['dout'])
        for a, b in [(out, g['out']), (x.grad, g['dx']), (c.grad, g['dcoeff']), (m.bias.grad, g['dbias'])]:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    elif name == "attention_case":
        layer = OracleDiffTransformerEncoderLayer(g['d'], g['H'], 2 * g['d'], 0.0)
        layer.zero_padded_queries = True
        layer.load_state_dict(g['state_dict'])
        src = g['src'].clone().requires_grad_()
        out, attn, heads = layer(src, pe=g['pe'], degree=g['degree'], src_key_padding_mask=g['mask'], need_heads=True)
        ((out * g['w']).sum() + (heads * g['wh']).sum()).backward()
        for a, b in [(out, g['out']), (attn, g['attn']), (heads, g['heads']), (src.grad, g['dsrc'])]:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    else:
        from feta_tmlr_b200 import synthetic
        m = synthetic.build_model("MUTAG", omodels, **g['over'])
        for layer in m.encoder.layers:
            layer.zero_padded_queries = True
        m.load_state_dict(g['state_dict'])
        px, mask, pe, lap, deg, labels, ei, bi, fi = g['batch']
        out, _, coeff = m(px, ei, bi, fi, mask, pe, lap, deg, return_filter_coeff=True)
        assert torch.allclose(out, g['out'], rtol=1e-5, atol=1e-6)
        assert torch.allclose(coeff, g['coeff'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("seed,sizes,self_loops", [(11, [5, 1, 7, 3], True), (12, [1, 1, 2], True), (13, [9, 4], False),
                                                   (14, [3, 6, 1, 1, 8], True)])
def test_arma_oracle_matches_dense_closed_form(seed, sizes, self_loops):
    """ARMAConvDynamic restatement (literal per-node weights + bmm) == dense closed form in fp64:
    mean_k relu(a_k A_hat x W_k + b_k x V_k + bias_k), A_hat = D_in^-1/2 A^T-accumulate D_in^-1/2 -- with isolated
    nodes (degree 0 -> 0, not inf), kept self-loops, multi-edges and directed extras."""
    from helpers import random_batch_graph
    from oracle.arma import arma_conv_dynamic
    ei, batch, R = random_batch_graph(seed, sizes, directed_extra=2, self_loops=self_loops)
    K, F, G = 3, 4, len(sizes)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(R, F, generator=g, dtype=torch.float64)
    coeff = torch.randn(G, 2 * K, generator=g, dtype=torch.float64)
    W = torch.randn(K, F, F, generator=g, dtype=torch.float64)
    V = torch.randn(1, K, F, F, generator=g, dtype=torch.float64)
    bias = torch.randn(1, K, 1, F, generator=g, dtype=torch.float64)
    out = arma_conv_dynamic(x, ei, coeff, batch.double(), W, None, V, bias, K)
    A = torch.zeros(R, R, dtype=torch.float64)
    for s, t in ei.t().tolist():
        A[t, s] += 1.0                                   # out[col] += x[row]
    deg = A.sum(1)                                       # degree over the target index
    dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
    Ah = dis[:, None] * A * dis[None, :]
    ref = torch.zeros(R, F, dtype=torch.float64)
    for k in range(K):
        a, b = coeff[batch, k][:, None], coeff[batch, K + k][:, None]
        ref += torch.relu(a * (Ah @ x @ W[k]) + b * (x @ V[0, k]) + bias[0, k])
    ref /= K
    assert out.shape == (R, F)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
