"""GPU parity: fused kernel-biased attention layer vs the CPU oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle.layers import OracleDiffTransformerEncoderLayer

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _inputs(seed, B, nmax, d, lens=None, with_pe=True):
    g = torch.Generator().manual_seed(seed)
    if lens is None:
        lens = torch.randint(1, nmax + 1, (B,), generator=g)
        lens[0] = nmax
    lens = torch.as_tensor(lens)
    mask = torch.arange(nmax)[None, :] >= lens[:, None]
    src = torch.randn(nmax, B, d, generator=g)
    pe = None
    if with_pe:
        a = torch.rand(B, nmax, nmax, generator=g)
        pe = (a + a.transpose(1, 2)) * 0.5
        pe = pe * (torch.rand(B, nmax, nmax, generator=g) > 0.2)          # exact zeros are meaningful
        valid = (~mask)[:, :, None] & (~mask)[:, None, :]
        pe = pe * valid                                                    # zero padded (data.py:182,212)
    degree = torch.rand(B, nmax, generator=g) * (~mask)
    return src, pe, degree, mask


def _pair(cuda, d, H, seed, **kw):
    from feta_tmlr_b200 import DiffTransformerEncoderLayer
    torch.manual_seed(seed)
    o = OracleDiffTransformerEncoderLayer(d, H, 2 * d, 0.0, **kw)
    o.zero_padded_queries = True                                           # product's documented convention
    m = DiffTransformerEncoderLayer(d, H, 2 * d, 0.0, **kw).to(cuda)
    m.load_state_dict(o.state_dict())
    return o, m


@pytest.fixture(params=["cuda_core", "tcgen05", "tcgen05_linear"])
def tensor_cores(request, monkeypatch):
    """Run the test through both attention forward kernels (fp32 CUDA-core / tcgen05 3xTF32), and with the layer's
    Linear layers on the tcgen05 path (csrc/linear_tc5.cu, fused Linear + residual + LayerNorm)."""
    from feta_tmlr_b200 import ops
    monkeypatch.setattr(ops, "ATTN_TENSOR_CORES", request.param == "tcgen05")
    monkeypatch.setattr(ops, "LINEAR_TC5", request.param == "tcgen05_linear")
    return request.param


@pytest.mark.parametrize("d,H", [(64, 8), (64, 4), (32, 1), (64, 1), (16, 4)])
@pytest.mark.parametrize("nmax", [7, 38, 100])
def test_attention_layer_parity(cuda, d, H, nmax, tensor_cores):
    o, m = _pair(cuda, d, H, seed=d + H + nmax)
    src, pe, degree, mask = _inputs(nmax, 5, nmax, d)
    so = src.clone().requires_grad_()
    oo, oa, oh = o(so, pe=pe, degree=degree, src_key_padding_mask=mask, need_heads=True)
    sg = src.to(cuda).requires_grad_()
    go, ga, gh = m(sg, pe=pe.to(cuda), degree=degree.to(cuda), src_key_padding_mask=mask.to(cuda), need_heads=True)
    assert ga.shape == oa.shape and gh.shape == oh.shape
    assert rel_err(ga, oa) < TOL and rel_err(gh, oh) < TOL and rel_err(go, oo) < TOL
    assert torch.equal(ga.cpu() == 0, oa == 0)                             # exact zeros preserved (models.py:276)
    w = torch.randn(oo.shape, generator=torch.Generator().manual_seed(1))
    wh = torch.randn(oh.shape, generator=torch.Generator().manual_seed(2))
    wa = torch.randn(oa.shape, generator=torch.Generator().manual_seed(3))
    ((oo * w).sum() + (oh * wh).sum() + (oa * wa).sum()).backward()
    ((go * w.to(cuda)).sum() + (gh * wh.to(cuda)).sum() + (ga * wa.to(cuda)).sum()).backward()
    assert rel_err(sg.grad, so.grad) < TOL
    for (n1, p1), (n2, p2) in zip(o.named_parameters(), m.named_parameters()):
        assert n1 == n2
        assert rel_err(p2.grad, p1.grad) < 2e-4, n1


@pytest.mark.parametrize("kw", [dict(), dict(share_qk=True), dict(attn_bias=True), dict(batch_norm=True)])
def test_attention_variants(cuda, kw):
    o, m = _pair(cuda, 32, 4, seed=11, **kw)
    src, pe, degree, mask = _inputs(3, 6, 20, 32)
    so, sg = src.clone().requires_grad_(), src.to(cuda).requires_grad_()
    oo, oa = o(so, pe=pe, degree=degree, src_key_padding_mask=mask)
    go, ga = m(sg, pe=pe.to(cuda), degree=degree.to(cuda), src_key_padding_mask=mask.to(cuda))
    assert rel_err(go, oo) < TOL and rel_err(ga, oa) < TOL
    w = torch.randn(oo.shape, generator=torch.Generator().manual_seed(1))   # (sum of squares of a LayerNorm
    (oo * w).sum().backward()                                                #  output has a ~zero gradient)
    (go * w.to(cuda)).sum().backward()
    assert rel_err(sg.grad, so.grad) < 2e-4
    for (n1, p1), (n2, p2) in zip(o.named_parameters(), m.named_parameters()):
        if float(p1.grad.abs().max()) < 1e-4:      # e.g. a bias feeding BatchNorm: true gradient is 0
            assert float((p2.grad.cpu() - p1.grad).abs().max()) < 1e-5, n1
        else:
            assert rel_err(p2.grad, p1.grad) < 5e-4, n1


def test_attention_no_pe_and_pe_diag_scaling(cuda, tensor_cores):
    o, m = _pair(cuda, 32, 4, seed=12)
    src, pe, degree, mask = _inputs(4, 4, 15, 32)
    oo, oa = o(src, pe=None, degree=degree, src_key_padding_mask=mask)     # pe=None: plain softmax rows
    go, ga = m(src.to(cuda), pe=None, degree=degree.to(cuda), src_key_padding_mask=mask.to(cuda))
    assert rel_err(go, oo) < TOL and rel_err(ga, oa) < TOL
    real = ~mask
    rows = ga.cpu().sum(-1)[real[:, None, :].expand(-1, 4, -1)]
    assert torch.allclose(rows, torch.ones_like(rows), atol=1e-5)
    pe2 = pe + torch.eye(15)[None] * (~mask)[:, :, None]
    oo, _ = o(src, pe=pe2, degree=None, src_key_padding_mask=mask)         # degree=None -> pe-diagonal scaling
    go, _ = m(src.to(cuda), pe=pe2.to(cuda), degree=None, src_key_padding_mask=mask.to(cuda))
    assert rel_err(go, oo) < TOL


@pytest.mark.parametrize("legacy", [False, True], ids=["tiled", "one_lds_per_fma"])
@pytest.mark.parametrize("d,with_pe", [(64, False), (32, True)])
def test_attention_large_graph_pattern_shape(cuda, tensor_cores, monkeypatch, legacy, d, with_pe):
    """Graphs of > 64 nodes go through the register-tiled forward / backward kernels (dh 16 and 8); the
    environment switches route the same shapes through the general kernels."""
    if legacy:
        monkeypatch.setenv("FETA_ATTN_FWD_LEGACY", "1")
        monkeypatch.setenv("FETA_ATTN_BWD_LEGACY", "1")
    o, m = _pair(cuda, d, 4, seed=13)
    src, pe, degree, mask = _inputs(5, 3, 188, d, lens=[188, 44, 120], with_pe=with_pe)
    if with_pe:
        pe = pe.clone()
    so, sg = src.clone().requires_grad_(), src.to(cuda).requires_grad_()
    oo, oa, oh = o(so, pe=pe, degree=degree, src_key_padding_mask=mask, need_heads=True)
    go, ga, gh = m(sg, pe=None if pe is None else pe.to(cuda), degree=degree.to(cuda),
                   src_key_padding_mask=mask.to(cuda), need_heads=True)
    assert rel_err(go, oo) < TOL and rel_err(ga, oa) < TOL and rel_err(gh, oh) < TOL
    (oo.sum() + oh.square().sum()).backward()
    (go.sum() + gh.square().sum()).backward()
    assert rel_err(sg.grad, so.grad) < TOL


def test_attention_rejects_unsupported(cuda):
    from feta_tmlr_b200 import DiffTransformerEncoderLayer
    m = DiffTransformerEncoderLayer(36, 3, 72, 0.0).to(cuda)               # head dim 12 not templated
    src, pe, degree, mask = _inputs(6, 2, 5, 36)
    with pytest.raises(RuntimeError):
        m(src.to(cuda), pe=pe.to(cuda), degree=degree.to(cuda), src_key_padding_mask=mask.to(cuda))
    m = DiffTransformerEncoderLayer(32, 4, 64, 0.0).to(cuda)
    src, pe, degree, mask = _inputs(6, 2, 5, 32)
    with pytest.raises(NotImplementedError):                               # attn_mask is never passed by the reference
        m(src.to(cuda), pe=pe.to(cuda), degree=degree.to(cuda), src_mask=torch.zeros(5, 5, device=cuda),
          src_key_padding_mask=mask.to(cuda))


@pytest.mark.parametrize("d,H,nmax", [(64, 8, 37), (64, 4, 150), (32, 1, 20)])
def test_attention_weight_dropout_with_injected_mask(cuda, monkeypatch, d, H, nmax):
    """--dropout > 0 (attention-weight dropout, F.dropout on P before P V): the kernels take the multipliers
    (0 or 1/(1-p)) as a tensor, so the oracle can be fed the SAME mask.  Forward, returned (dropped) attention and
    every gradient, including the gradient flowing through the returned attention matrix."""
    import oracle.layers as olayers
    from feta_tmlr_b200.layers import DiffMultiheadAttention
    p = 0.25
    B = 4
    torch.manual_seed(d + nmax)
    o = olayers.OracleDiffMultiheadAttention(d, H, dropout=p)
    m = DiffMultiheadAttention(d, H, dropout=p).to(cuda)
    m.load_state_dict(o.state_dict())
    o.train(), m.train()
    src, pe, degree, mask = _inputs(nmax + 1, B, nmax, d)
    g = torch.Generator().manual_seed(7)
    dm = (torch.rand(B, H, nmax, nmax, generator=g) >= p).float() / (1.0 - p)
    monkeypatch.setattr(olayers.F, "dropout", lambda w, p=0.5, training=True: w * dm if w.dim() == 4 else w)
    so = src.clone().requires_grad_()
    oo, oa, oh = o(so, pe=pe, key_padding_mask=mask, zero_padded_queries=True)
    sg = src.to(cuda).requires_grad_()
    go, ga, gh = m(sg, pe=pe.to(cuda), key_padding_mask=mask.to(cuda), drop=dm.to(cuda))
    assert rel_err(go, oo) < TOL and rel_err(ga, oa) < TOL and rel_err(gh, oh) < TOL
    w = torch.randn(oo.shape, generator=g)
    wa = torch.randn(oa.shape, generator=g)
    ((oo * w).sum() + (oa * wa).sum()).backward()
    ((go * w.to(cuda)).sum() + (ga * wa.to(cuda)).sum()).backward()
    assert rel_err(sg.grad, so.grad) < TOL
    for (n1, p1), (n2, p2) in zip(o.named_parameters(), m.named_parameters()):
        assert rel_err(p2.grad, p1.grad) < 2e-4, n1


def test_attention_dropout_trains_and_is_off_in_eval(cuda):
    """The layer draws its own multipliers in training mode (no NotImplementedError any more) and is deterministic
    in eval mode."""
    from feta_tmlr_b200 import DiffTransformerEncoderLayer
    torch.manual_seed(0)
    m = DiffTransformerEncoderLayer(64, 8, 128, 0.2).to(cuda)
    src, pe, degree, mask = _inputs(3, 6, 30, 64)
    args = dict(pe=pe.to(cuda), degree=degree.to(cuda), src_key_padding_mask=mask.to(cuda))
    m.train()
    s = src.to(cuda).requires_grad_()
    out, attn = m(s, **args)
    out.square().mean().backward()
    assert torch.isfinite(s.grad).all() and float(s.grad.abs().max()) > 0
    frac_zero = float(((attn == 0) & (~mask.to(cuda))[:, None, :, None] & (~mask.to(cuda))[:, None, None, :]).float().mean())
    assert frac_zero > 0.02                                   # some real weights were dropped
    m.eval()
    with torch.no_grad():
        a, _ = m(src.to(cuda), **args)
        b, _ = m(src.to(cuda), **args)
    assert torch.equal(a, b)


# ---- matrix-free kernels (csrc/attention_rows.cu): need_attn=False layers --------------------------------------
@pytest.mark.parametrize("d,H", [(64, 8), (64, 4), (32, 1), (16, 4), (32, 8)])
@pytest.mark.parametrize("nmax,with_pe", [(7, True), (37, True), (38, False), (100, True), (188, False), (256, True)])
def test_attention_rows_layer_parity(cuda, d, H, nmax, with_pe):
    """``need_attn=False``: no attention matrix is written; output, heads and every gradient still equal the oracle
    layer's (whose matrix is simply not used)."""
    from feta_tmlr_b200 import ops
    assert ops.attn_rows_enabled(nmax, d // H)
    o, m = _pair(cuda, d, H, seed=d + H + nmax)
    src, pe, degree, mask = _inputs(nmax + 1, 4, nmax, d, with_pe=with_pe)
    so = src.clone().requires_grad_()
    oo, _, oh = o(so, pe=pe, degree=degree, src_key_padding_mask=mask, need_heads=True)
    sg = src.to(cuda).requires_grad_()
    go, ga, gh = m(sg, pe=None if pe is None else pe.to(cuda), degree=degree.to(cuda),
                   src_key_padding_mask=mask.to(cuda), need_heads=True, need_attn=False)
    assert ga is None
    assert rel_err(gh, oh) < TOL and rel_err(go, oo) < TOL
    w = torch.randn(oo.shape, generator=torch.Generator().manual_seed(1))
    wh = torch.randn(oh.shape, generator=torch.Generator().manual_seed(2))
    ((oo * w).sum() + (oh * wh).sum()).backward()
    ((go * w.to(cuda)).sum() + (gh * wh.to(cuda)).sum()).backward()
    assert rel_err(sg.grad, so.grad) < TOL
    for (n1, p1), (n2, p2) in zip(o.named_parameters(), m.named_parameters()):
        assert rel_err(p2.grad, p1.grad) < 2e-4, n1


def _attn_core_oracle(qkv, pe, mask, H, scale):
    """fp64 restatement of the attention core (oracle/layers.py: scores, key mask, max, exp * pe, clamp, P V) with
    the product's zero-row convention for masked queries."""
    N, B, d3 = qkv.shape
    d, dh = d3 // 3, d3 // 3 // H
    q, k, v = [t.reshape(N, B, H, dh).permute(1, 2, 0, 3) for t in qkv.split(d, dim=-1)]      # [B,H,N,dh]
    s = (q * scale) @ k.transpose(-1, -2)
    s = s.masked_fill(mask[:, None, None, :], float('-inf'))
    mx = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - torch.where(torch.isinf(mx), torch.zeros_like(mx), mx))      # a graph without a real key: e = 0
    if pe is not None:
        e = e * pe[:, None]
    p = e / e.sum(dim=-1, keepdim=True).clamp(min=1e-6)
    o = (p @ v) * (~mask)[:, None, :, None]
    return o.permute(2, 0, 1, 3)                                                              # [N,B,H,dh]


@pytest.mark.parametrize("case", ["interior_mask", "all_masked_graph", "clamped_rows", "share_qk"])
def test_attention_rows_core_edge_cases(cuda, case):
    """The core through ops.diff_attention(need_attn=False): a padding mask that is not a suffix (generic loops), a
    graph with no real node, rows whose kernel-weighted sum falls under the 1e-6 clamp, shared q/k projections."""
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(7)
    B, N, H, dh = 3, 45, 4, 8
    d = H * dh
    qkv = torch.randn(N, B, 3 * d, generator=g, dtype=torch.float64)
    lens = torch.tensor([45, 20, 33])
    mask = torch.arange(N)[None, :] >= lens[:, None]
    a = torch.rand(B, N, N, generator=g, dtype=torch.float64)
    pe = (a + a.transpose(1, 2)) * 0.5 * (torch.rand(B, N, N, generator=g) > 0.3)
    share = case == "share_qk"
    if case == "interior_mask":
        mask[0, 3] = True
        mask[0, 17] = True
        mask[2, 0] = True
    if case == "all_masked_graph":
        mask[1, :] = True
    if case == "clamped_rows":
        pe[0, 5, :] = 0.0                          # sum 0: P = 0 / 1e-6
        pe[2, 7, :] *= 1e-9                        # sum under the clamp: P = e * pe / 1e-6, no delta term
    pe = pe * ((~mask)[:, :, None] & (~mask)[:, None, :])
    if share:
        qkv[..., d:2 * d] = qkv[..., :d]
    scale = dh ** -0.5
    ref_in = qkv.clone().requires_grad_()
    if share:
        qq = ref_in[..., :d]
        o_ref = _attn_core_oracle(torch.cat([qq, qq, ref_in[..., 2 * d:]], dim=-1), pe, mask, H, scale)
    else:
        o_ref = _attn_core_oracle(ref_in, pe, mask, H, scale)
    w = torch.randn(o_ref.shape, generator=g, dtype=torch.float64)
    (o_ref * w).sum().backward()
    x = qkv.float().to(cuda).requires_grad_()
    attn, o = ops.diff_attention(x, pe.float().to(cuda), mask.to(cuda), H, scale, share_qk=share, need_attn=False)
    assert attn is None
    assert rel_err(o, o_ref.float()) < TOL
    (o * w.float().to(cuda)).sum().backward()
    gref = ref_in.grad.float()
    if share:
        gref = gref.clone()
        gref[..., d:2 * d] = 0
    assert rel_err(x.grad, gref) < TOL
    # the matrix-writing kernels agree on the same inputs
    x2 = qkv.float().to(cuda).requires_grad_()
    _, o2 = ops.diff_attention(x2, pe.float().to(cuda), mask.to(cuda), H, scale, share_qk=share, need_attn=True)
    (o2 * w.float().to(cuda)).sum().backward()
    assert rel_err(o, o2) < 1e-5 and rel_err(x.grad, x2.grad) < 1e-4


def test_attention_rows_switch_off(cuda, monkeypatch):
    from feta_tmlr_b200 import ops
    monkeypatch.setattr(ops, "ATTN_ROWS", False)
    o, m = _pair(cuda, 32, 4, seed=3)
    src, pe, degree, mask = _inputs(3, 3, 12, 32)
    go, ga = m(src.to(cuda), pe=pe.to(cuda), degree=degree.to(cuda), src_key_padding_mask=mask.to(cuda),
               need_attn=False)
    assert ga is not None                                   # falls back to the matrix-writing kernels


@pytest.mark.parametrize("B,N,H,dh,with_pe", [(4, 37, 8, 8, True), (3, 150, 4, 16, False), (5, 20, 2, 32, True),
                                               (2, 256, 1, 16, True)])
@pytest.mark.parametrize("layout", ["packed", "padded_slots", "interior_mask"])
def test_lazy_attention_coeff_scalar_matches_matrix_path(cuda, B, N, H, dh, with_pe, layout):
    """N1: the coefficient scalar recomputed from q / k (feta_attn_rows_coeff) == feta_coeff_scalar over the
    materialised attention matrix -- packed output (reference layout), padded slots (static layout), a mask that is
    not a suffix, exact zeros on the diagonal of the kernel (loop weight 1, PyG add_remaining_self_loops)."""
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(B * N + H)
    d = H * dh
    qkv = torch.randn(N, B, 3 * d, generator=g).to(cuda)
    lens = torch.randint(1, N + 1, (B,), generator=g)
    lens[0] = N
    mask = torch.arange(N)[None, :] >= lens[:, None]
    if layout == "interior_mask":
        mask[0, 1] = True
        mask[B - 1, 0] = True
    pe = None
    if with_pe:
        a = torch.rand(B, N, N, generator=g)
        pe = (a + a.transpose(1, 2)) * 0.5 * (torch.rand(B, N, N, generator=g) > 0.2)
        idx = torch.arange(0, N, 3)
        pe[:, idx, idx] = 0.0                                      # zero self weights: loop weight falls back to 1
        pe = (pe * ((~mask)[:, :, None] & (~mask)[:, None, :])).to(cuda)
    mask = mask.to(cuda)
    scale = dh ** -0.5
    lazy, o1 = ops.diff_attention(qkv, pe, mask, H, scale, need_attn='coeff')
    assert isinstance(lazy, ops.LazyAttention)
    attn, o2 = ops.diff_attention(qkv, pe, mask, H, scale, need_attn=True)
    assert rel_err(o1, o2) < 1e-5
    assert rel_err(lazy.materialize(), attn) == 0.0
    real = (~mask).sum(dim=1)
    if layout == "padded_slots":
        node_ptr = (torch.arange(B, device=cuda) * N).to(torch.int32)
        total, zero_fill = B * N, True
    else:
        node_ptr = (torch.cumsum(real, 0) - real).to(torch.int32)
        total, zero_fill = int(real.sum()), False
    s_lazy = ops.coeff_scalar(lazy, mask, node_ptr, total, zero_fill=zero_fill)
    s_mat = ops.coeff_scalar(attn, mask, node_ptr, total, zero_fill=zero_fill)
    assert s_lazy.shape == s_mat.shape
    assert rel_err(s_lazy, s_mat) < 1e-5
