"""GPU parity of the layer-glue kernels (csrc/dense.cu): weight-gradient reduction, residual-add LayerNorm and
the residual-folding Linear, against plain PyTorch (fp64 on the CPU) -- with the parameter gradients on the side
stream (default) and on the launching stream."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(params=[True, False], ids=["side_stream", "main_stream"])
def side(request):
    from feta_tmlr_b200 import ops
    old = ops.WGRAD_SIDE_STREAM
    ops.WGRAD_SIDE_STREAM = request.param
    yield request.param
    ops.WGRAD_SIDE_STREAM = old


@pytest.mark.parametrize("T,D", [(1, 16), (37, 64), (4352, 64), (10752, 128), (300, 200), (5000, 256)])
@pytest.mark.parametrize("scaled", [False, True])
def test_add_layer_norm_parity(cuda, side, T, D, scaled):
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(T + D)
    a, b = torch.randn(T, D, generator=g), torch.randn(T, D, generator=g)
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    bs = torch.rand(T, generator=g) + 0.5 if scaled else None
    go = torch.randn(T, D, generator=g)
    ref_in = [t.double().requires_grad_() for t in (a, b, gamma, beta)]
    z = ref_in[0] + (ref_in[1] if bs is None else bs.double().unsqueeze(1) * ref_in[1])
    ref = torch.nn.functional.layer_norm(z, (D,), ref_in[2], ref_in[3], 1e-5)
    ref.backward(go.double())
    dev_in = [t.to(cuda).requires_grad_() for t in (a, b, gamma, beta)]
    out = ops.add_layer_norm(dev_in[0], dev_in[1], dev_in[2], dev_in[3], 1e-5,
                             bscale=None if bs is None else bs.to(cuda))
    out.backward(go.to(cuda))
    torch.cuda.synchronize()
    assert rel_err(out, ref) <= TOL
    for d, r in zip(dev_in, ref_in):
        assert rel_err(d.grad, r.grad) <= TOL


@pytest.mark.parametrize("T,fin,fout", [(1, 8, 8), (4352, 64, 192), (4352, 128, 64), (10752, 64, 128), (777, 20, 36)])
@pytest.mark.parametrize("relu,with_res", [(False, False), (True, True), (False, True)])
def test_linear_parity(cuda, side, tc, T, fin, fout, relu, with_res):
    """y = x W^T + b (optionally ReLU in the epilogue), dW / db through feta_linear_wgrad, and the residual
    output whose gradient is folded into the dX GEMM."""
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(T + fin)
    x = torch.randn(T, fin, generator=g)
    W, b = torch.randn(fout, fin, generator=g) * 0.2, torch.randn(fout, generator=g)
    go, gr = torch.randn(T, fout, generator=g), torch.randn(T, fin, generator=g)
    xr, Wr, br = (t.double().requires_grad_() for t in (x, W, b))
    ref = torch.nn.functional.linear(xr, Wr, br)
    if relu:
        ref = torch.relu(ref)
    ((ref * go.double()).sum() + ((xr * gr.double()).sum() if with_res else 0)).backward()
    xd, Wd, bd = (t.to(cuda).requires_grad_() for t in (x, W, b))
    if with_res:
        y, res = ops.linear_res(xd, Wd, bd, relu=relu)
        assert res.data_ptr() == xd.data_ptr()
        ((y * go.to(cuda)).sum() + (res * gr.to(cuda)).sum()).backward()
    else:
        y = ops.linear(xd, Wd, bd, relu=relu)
        (y * go.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(y, ref) <= TOL
    assert rel_err(xd.grad, xr.grad) <= TOL
    assert rel_err(Wd.grad, Wr.grad) <= TOL
    assert rel_err(bd.grad, br.grad) <= TOL


@pytest.fixture(params=["simt", "tcgen05", "mma_sync", "library_gemm"])
def tc(request, monkeypatch):
    """simt: csrc/linear_simt.cu (the default: exact-fp32 CUDA-core latency kernel; shapes it does not take fall back
    to the library);  tcgen05: csrc/linear_tc5.cu where the shape is a multiple of its tiles (else mma.sync);
    mma_sync: csrc/dense_tc.cu only;  library_gemm: torch / cuBLAS."""
    from feta_tmlr_b200 import ops
    monkeypatch.setattr(ops, "LINEAR_SIMT", request.param == "simt")
    monkeypatch.setattr(ops, "LINEAR_TC5", request.param == "tcgen05")
    monkeypatch.setattr(ops, "LINEAR_TENSOR_CORES", request.param in ("tcgen05", "mma_sync"))
    yield request.param != "library_gemm"


@pytest.mark.parametrize("T", [1, 31, 33, 4736, 12032])
@pytest.mark.parametrize("fin,fout", [(64, 64), (64, 192), (64, 128), (128, 64), (192, 64), (256, 256)])
def test_linear_simt_kernel(cuda, T, fin, fout):
    """csrc/linear_simt.cu through the C ABI (impl = SIMT): forward with bias + ReLU, dX with ReLU mask and residual,
    against fp64 -- exact-fp32 arithmetic, so the tolerance is fp32 rounding of a <= 256-term sum."""
    from feta_tmlr_b200 import ops, _lib
    lib = _lib.load()
    assert lib.feta_linear_simt_supported(fin, fout)
    g = torch.Generator().manual_seed(T * 7 + fin + fout)
    x = torch.randn(T, fin, generator=g).to(cuda)
    W = (torch.randn(fout, fin, generator=g) * 0.2).to(cuda)
    b = torch.randn(fout, generator=g).to(cuda)
    dy = torch.randn(T, fout, generator=g).to(cuda)
    dres = torch.randn(T, fin, generator=g).to(cuda)
    msk = torch.randn(T, fin, generator=g).to(cuda)
    y = torch.empty(T, fout, device=cuda)
    dx = torch.empty(T, fin, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    P = ops._ptr
    for relu in (0, 1):
        ops.check(lib.feta_linear_fwd_ex(P(x), P(W), P(b), P(y), T, fin, fout, relu, 1, st), "fwd")
        ref = x.double() @ W.double().t() + b.double()
        if relu:
            ref = torch.relu(ref)
        assert rel_err(y, ref) < 2e-6
    ops.check(lib.feta_linear_fwd_ex(P(x), P(W), None, P(y), T, fin, fout, 0, 1, st), "fwd no bias")
    assert rel_err(y, x.double() @ W.double().t()) < 2e-6
    ops.check(lib.feta_linear_dx_ex(P(dy), P(W), P(dres), P(msk), P(dx), T, fin, fout, 1, st), "dx")
    ref = (dy.double() @ W.double()) * (msk > 0).double() + dres.double()
    assert rel_err(dx, ref) < 2e-6
    ops.check(lib.feta_linear_dx_ex(P(dy), P(W), None, None, P(dx), T, fin, fout, 1, st), "dx plain")
    assert rel_err(dx, dy.double() @ W.double()) < 2e-6


@pytest.mark.parametrize("T,d,dff", [(1, 8, 16), (4352, 64, 128), (10752, 64, 128), (37, 128, 256), (90, 24, 40)])
def test_ffn_chain_parity(cuda, tc, T, d, dff):
    """linear1 (+ReLU epilogue, gradient pre-masked) -> linear2 (ReLU mask in its dX epilogue) with the residual
    routed through linear1: forward and every gradient against fp64 PyTorch."""
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(T + d)
    x = torch.randn(T, d, generator=g)
    W1, b1 = torch.randn(dff, d, generator=g) * 0.2, torch.randn(dff, generator=g) * 0.1
    W2, b2 = torch.randn(d, dff, generator=g) * 0.2, torch.randn(d, generator=g) * 0.1
    go = torch.randn(T, d, generator=g)
    ref_in = [t.double().requires_grad_() for t in (x, W1, b1, W2, b2)]
    ref = ref_in[0] + torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(ref_in[0], ref_in[1], ref_in[2])),
                                                  ref_in[3], ref_in[4])
    ref.backward(go.double())
    dv = [t.to(cuda).requires_grad_() for t in (x, W1, b1, W2, b2)]
    h, res = ops.linear_res(dv[0], dv[1], dv[2], relu=True, grad_premasked=True)
    out = res + ops.linear(h, dv[3], dv[4], mask_input_grad=True)
    out.backward(go.to(cuda))
    torch.cuda.synchronize()
    assert rel_err(out, ref) <= TOL
    for a, r in zip(dv, ref_in):
        assert rel_err(a.grad, r.grad) <= TOL


def test_linear_tc_accuracy_is_fp32_grade(cuda):
    """3xTF32 (hi.hi + hi.lo + lo.hi) must sit at fp32 round-off, far below plain TF32's ~1e-3."""
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(0)
    x, W = torch.randn(4352, 64, generator=g), torch.randn(192, 64, generator=g)
    ref = x.double() @ W.double().t()
    y = ops.linear(x.to(cuda), W.to(cuda))
    assert rel_err(y, ref) < 2e-6


@pytest.mark.parametrize("weight_decay", [0.0, 0.05])
def test_flat_adam_matches_torch(cuda, weight_decay):
    """feta_adam_step over flat buffers == torch.optim.Adam / AdamW, parameter by parameter, over several steps
    (odd sizes: the flat layout is packed, the kernel's float4 body + scalar tail must cover everything)."""
    from feta_tmlr_b200 import ddp, engine
    g = torch.Generator().manual_seed(3)
    shapes = [(7, 5), (13,), (64, 64), (1,), (3, 3, 3), (130,)]
    ref = [torch.randn(*s, generator=g).to(cuda).requires_grad_() for s in shapes]
    mine = [p.detach().clone().requires_grad_() for p in ref]
    cls = torch.optim.AdamW if weight_decay else torch.optim.Adam
    kw = dict(weight_decay=weight_decay) if weight_decay else {}
    opt = cls(ref, lr=3e-3, **kw)
    bucket = ddp.FlatGradBucket(mine, attach=False)
    fa = engine.FlatAdam(bucket, lr=3e-3, weight_decay=weight_decay)
    assert all(p.data_ptr() >= fa.flat_p.data_ptr() for p in mine)         # parameters now live in the flat buffer
    for step in range(6):
        grads = [torch.randn(*s, generator=g).to(cuda) * (0.1 + step) for s in shapes]
        if step == 3:
            for pg in opt.param_groups:
                pg['lr'] = 1e-3
            fa.set_lr(1e-3)
        for p, gr in zip(ref, grads):
            p.grad = gr.clone()
        opt.step()
        torch._foreach_copy_(bucket.views, grads)
        fa.step(grad_scale=1.0)
        for a, b in zip(mine, ref):
            assert rel_err(a, b) < 2e-6, step
    assert float(fa.step_count) == 6.0


def test_gradient_accumulation_over_two_backward_passes(cuda):
    """Default mode (no opt-in side stream): a second backward pass accumulates into existing ``.grad`` tensors --
    the case the side-stream variant must never be used for (autograd adds on the main stream)."""
    from feta_tmlr_b200 import ops
    from feta_tmlr_b200.layers import DiffTransformerEncoderLayer
    assert ops.WGRAD_SIDE_STREAM is False
    torch.manual_seed(3)
    d, H, B, nmax = 64, 4, 48, 40
    layer = DiffTransformerEncoderLayer(d, H, 2 * d, 0.0).to(cuda)
    src = torch.randn(nmax, B, d, device=cuda)
    lens = torch.randint(5, nmax + 1, (B,), device=cuda)
    mask = torch.arange(nmax, device=cuda)[None, :] >= lens[:, None]
    pe = torch.rand(B, nmax, nmax, device=cuda)
    deg = torch.rand(B, nmax, device=cuda)

    def run():
        out, _ = layer(src, pe=pe, degree=deg, src_key_padding_mask=mask)
        out.square().mean().backward()

    run()
    torch.cuda.synchronize()
    once = {k: p.grad.clone() for k, p in layer.named_parameters()}
    for _ in range(3):
        run()                                   # accumulates: grad = 4 x once
    torch.cuda.synchronize()
    for k, p in layer.named_parameters():
        assert rel_err(p.grad, 4 * once[k]) < 1e-5, k


def test_side_stream_opt_in_matches_main_stream(cuda):
    from feta_tmlr_b200 import ops
    from feta_tmlr_b200.layers import DiffTransformerEncoderLayer
    torch.manual_seed(4)
    d, H, B, nmax = 64, 8, 32, 30
    layer = DiffTransformerEncoderLayer(d, H, 2 * d, 0.0).to(cuda)
    src = torch.randn(nmax, B, d, device=cuda)
    pe = torch.rand(B, nmax, nmax, device=cuda)
    deg = torch.rand(B, nmax, device=cuda)
    grads = []
    for on in (False, True):
        for p in layer.parameters():
            p.grad = None
        out, _ = layer(src, pe=pe, degree=deg)
        with ops.wgrad_side_stream(on):
            out.square().mean().backward()
        torch.cuda.synchronize()
        grads.append({k: p.grad.clone() for k, p in layer.named_parameters()})
    assert ops.WGRAD_SIDE_STREAM is False
    for k in grads[0]:
        assert torch.equal(grads[0][k], grads[1][k]), k


@pytest.mark.parametrize("T,D", [(37, 64), (4736, 64), (1000, 16), (777, 128)])
@pytest.mark.parametrize("scaled,weighted", [(False, False), (True, False), (True, True)])
def test_add_batch_norm_parity(cuda, T, D, scaled, weighted):
    """Fused residual + BatchNorm1d (training) vs torch.nn.BatchNorm1d in fp64: output, every gradient, running
    statistics (unbiased variance, num_batches_tracked); with row weights the statistics come from the kept rows."""
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(T + D)
    a, b = torch.randn(T, D, generator=g) + 0.5, torch.randn(T, D, generator=g)
    bs = torch.rand(T, generator=g) + 0.5 if scaled else None
    w = (torch.rand(T, generator=g) < 0.7).float() if weighted else None
    go = torch.randn(T, D, generator=g)
    bn_ref = torch.nn.BatchNorm1d(D).double()
    bn_ref.weight.data.uniform_(0.5, 1.5, generator=g)
    bn_ref.bias.data.uniform_(-0.5, 0.5, generator=g)
    bn = torch.nn.BatchNorm1d(D)
    bn.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in bn_ref.state_dict().items()})
    bn = bn.to(cuda).train()
    ar, br = a.double().requires_grad_(), b.double().requires_grad_()
    z = ar + (br if bs is None else bs.double().unsqueeze(1) * br)
    keep = torch.ones(T, dtype=torch.bool) if w is None else w.bool()
    yk = bn_ref(z[keep])
    (yk * go.double()[keep]).sum().backward()
    ad, bd = a.to(cuda).requires_grad_(), b.to(cuda).requires_grad_()
    y = ops.add_batch_norm(ad, bd, bn, bscale=None if bs is None else bs.to(cuda), roww=None if w is None else w.to(cuda))
    (y * go.to(cuda))[keep.to(cuda)].sum().backward()
    torch.cuda.synchronize()
    assert rel_err(y[keep.to(cuda)], yk) < TOL
    assert rel_err(ad.grad, ar.grad) < TOL and rel_err(bd.grad, br.grad) < TOL
    assert rel_err(bn.weight.grad, bn_ref.weight.grad) < TOL and rel_err(bn.bias.grad, bn_ref.bias.grad) < TOL
    assert rel_err(bn.running_mean, bn_ref.running_mean) < 1e-5 and rel_err(bn.running_var, bn_ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize("T", [1, 127, 128, 129, 4736, 12032])
@pytest.mark.parametrize("fin,fout", [(64, 192), (64, 64), (64, 128), (128, 64), (192, 64), (256, 128)])
def test_linear_tcgen05_only_switch(cuda, monkeypatch, T, fin, fout):
    """FETA_LINEAR_TC5 alone (library GEMM elsewhere): forward with bias + ReLU, dX with ReLU mask and residual."""
    from feta_tmlr_b200 import ops
    monkeypatch.setattr(ops, "LINEAR_TC5", True)
    monkeypatch.setattr(ops, "LINEAR_TENSOR_CORES", False)
    assert ops.linear_tc_enabled(fin, fout)
    g = torch.Generator().manual_seed(T + fin + fout)
    x = torch.randn(T, fin, generator=g)
    W, b = torch.randn(fout, fin, generator=g) * 0.2, torch.randn(fout, generator=g)
    go, gr = torch.randn(T, fout, generator=g), torch.randn(T, fin, generator=g)
    xd, Wd, bd = (t.to(cuda).requires_grad_() for t in (x, W, b))
    y, res = ops.linear_res(xd, Wd, bd, relu=True)
    xr, Wr, br = (t.double().requires_grad_() for t in (x, W, b))
    # the ReLU mask of the fp64 reference is taken from the kernel's output: a pre-activation within fp32 rounding of
    # zero may legitimately fall on either side (1.5 M outputs at the largest shape)
    ref = torch.nn.functional.linear(xr, Wr, br) * (y.detach().cpu() > 0).double()
    ((ref * go.double()).sum() + (xr * gr.double()).sum()).backward()
    ((y * go.to(cuda)).sum() + (res * gr.to(cuda)).sum()).backward()
    torch.cuda.synchronize()
    assert rel_err(y, ref) < TOL
    assert rel_err(xd.grad, xr.grad) < TOL and rel_err(Wd.grad, Wr.grad) < TOL and rel_err(bd.grad, br.grad) < TOL


@pytest.mark.parametrize("T", [1, 130, 4736])
@pytest.mark.parametrize("fin", [64, 128])
@pytest.mark.parametrize("scaled,masked", [(True, False), (False, True)])
def test_linear_add_layer_norm_fused(cuda, side, T, fin, scaled, masked):
    """One-launch [Linear + degree scale + residual + LayerNorm] (tcgen05) vs fp64 PyTorch: y and every gradient;
    `masked`: the Linear's input is a ReLU output whose mask is applied in the dX epilogue."""
    from feta_tmlr_b200 import ops
    D = 64
    g = torch.Generator().manual_seed(T + fin)
    x = torch.randn(T, fin, generator=g)
    if masked:
        x = torch.relu(x)
    W, b = torch.randn(D, fin, generator=g) * 0.2, torch.randn(D, generator=g)
    res = torch.randn(T, D, generator=g)
    gamma, beta = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g)
    bs = torch.rand(T, generator=g) + 0.5 if scaled else None
    go = torch.randn(T, D, generator=g)
    pre = torch.randn(T, fin, generator=g)                     # pre-activation whose ReLU is x (masked case)
    xr, Wr, br, rr, gr, ber = (t.double().requires_grad_() for t in (x, W, b, res, gamma, beta))
    lin = torch.nn.functional.linear(xr, Wr, br)
    z = rr + (lin if bs is None else bs.double().unsqueeze(1) * lin)
    ref = torch.nn.functional.layer_norm(z, (D,), gr, ber, 1e-5)
    ref.backward(go.double())
    xd, Wd, bd, rd, gd, bed = (t.to(cuda).requires_grad_() for t in (x, W, b, res, gamma, beta))
    y = ops.linear_add_layer_norm(xd, Wd, bd, rd, gd, bed, 1e-5, bscale=None if bs is None else bs.to(cuda),
                                  mask_input_grad=masked)
    y.backward(go.to(cuda))
    torch.cuda.synchronize()
    assert rel_err(y, ref) < TOL
    want_dx = xr.grad * (x > 0).double() if masked else xr.grad
    assert rel_err(xd.grad, want_dx) < TOL
    for d_, r_ in ((Wd, Wr), (bd, br), (rd, rr), (gd, gr), (bed, ber)):
        assert rel_err(d_.grad, r_.grad) < TOL
