"""GPU parity: fused ARMAConvDynamic (csrc/arma.cu, NORM_GCN plan) vs the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import random_batch_graph, rel_err
from oracle.arma import arma_conv_dynamic
from oracle.pyg17 import gcn_norm

pytestmark = pytest.mark.gpu
TOL = 1e-4      # north_star: <= 1e-4 relative, fp32


@pytest.mark.parametrize("sizes", [[5, 7, 1, 4], [33, 2, 64, 17, 1, 1, 90], [200, 150]])
def test_gcn_plan_bit_exact(cuda, sizes):
    """gcn_norm(add_self_loops=False): self-loops kept, degree over the target index."""
    from feta_tmlr_b200 import ops
    ei, batch, R = random_batch_graph(3, sizes, directed_extra=2)
    p = ops.build_cheb_plan(ei.to(cuda), batch.to(cuda), R, len(sizes), 2.0, norm=ops.NORM_GCN)
    _, w = gcn_norm(ei, None, R, add_loops=False, dtype=torch.float32)
    s, t = ei[0].numpy(), ei[1].numpy()
    for key, other, (rp, ci, va) in ((t, s, (p.rowptr, p.colidx, p.vals)), (s, t, (p.rowptr_t, p.colidx_t, p.vals_t))):
        order = np.argsort(key, kind='stable')
        orp = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=R))]).astype(np.int32)
        nnz = int(orp[-1])
        assert nnz == ei.shape[1]
        assert np.array_equal(rp.cpu().numpy(), orp)
        assert np.array_equal(ci.cpu().numpy()[:nnz], other[order].astype(np.int32))
        np.testing.assert_allclose(va.cpu().numpy()[:nnz], w.numpy()[order], rtol=2e-7, atol=0)


def _case(seed, sizes, F, K, dropout_root=False, directed_extra=1):
    ei, batch, R = random_batch_graph(seed, sizes, directed_extra=directed_extra)
    G = len(sizes)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, F, generator=g)
    coeff = torch.randn(G, 2 * K, generator=g)
    xr = None
    if dropout_root:
        keep = (torch.rand(R, F, generator=g) > 0.3).float()
        xr = x * keep / 0.7
    return ei, batch, R, G, x, coeff, xr


@pytest.mark.parametrize("sizes,F,K", [([5, 7, 1, 4], 4, 2), ([9, 23, 37, 12, 1], 8, 4), ([33, 2, 64, 17, 1, 1, 90], 16, 4),
                                         ([188, 44, 120], 16, 3), ([222, 3], 16, 4)])
@pytest.mark.parametrize("float_batch", [False, True])
def test_arma_forward_backward_parity(cuda, sizes, F, K, float_batch):
    from feta_tmlr_b200 import ARMAConvDynamic
    ei, batch, R, G, x, coeff, _ = _case(5, sizes, F, K)
    torch.manual_seed(0)
    mod = ARMAConvDynamic(F, F, num_stacks=K, num_layers=1)
    mod.bias.data.normal_(0, 0.3)
    go = torch.randn(R, F, generator=torch.Generator().manual_seed(9))
    bt = batch.float() if float_batch else batch

    xo, co = x.clone().requires_grad_(), coeff.clone().requires_grad_()
    params = [p.detach().clone().requires_grad_() for p in (mod.init_weight, mod.root_weight, mod.bias)]
    ref = arma_conv_dynamic(xo, ei, co, bt, params[0], None, params[1], params[2], K)
    ref.backward(go)

    dm = mod.to(cuda)
    xd, cd = x.to(cuda).requires_grad_(), coeff.to(cuda).requires_grad_()
    out = dm(xd, ei.to(cuda), cd, batch=bt.to(cuda))
    out.backward(go.to(cuda))
    assert out.shape == ref.shape
    assert rel_err(out, ref) <= TOL
    assert rel_err(xd.grad, xo.grad) <= TOL
    assert rel_err(cd.grad, co.grad) <= TOL
    assert rel_err(dm.init_weight.grad, params[0].grad) <= TOL
    assert rel_err(dm.root_weight.grad, params[1].grad) <= TOL
    assert rel_err(dm.bias.grad, params[2].grad) <= TOL
    assert dm.weight.grad is None                          # unused when num_layers == 1, as in the reference


def test_arma_separate_root_and_determinism(cuda):
    """x_root != x (the dropout-ed skip input of :335) gets its own gradient; two runs are bit-identical."""
    from feta_tmlr_b200 import ops, ARMAConvDynamic
    sizes, F, K = [9, 23, 37, 12, 1], 8, 4
    ei, batch, R, G, x, coeff, xr = _case(7, sizes, F, K, dropout_root=True)
    torch.manual_seed(1)
    mod = ARMAConvDynamic(F, F, num_stacks=K)
    xo, ro = x.clone().requires_grad_(), xr.clone().requires_grad_()
    ref = arma_conv_dynamic(xo, ei, coeff, batch, mod.init_weight, None, mod.root_weight, mod.bias, K, x_root=ro)
    ref.square().sum().backward()
    dm = mod.to(cuda)
    plan = ops.build_cheb_plan(ei.to(cuda), batch.to(cuda), R, G, 2.0, norm=ops.NORM_GCN)
    outs = []
    for _ in range(2):
        xd, rd = x.to(cuda).requires_grad_(), xr.to(cuda).requires_grad_()
        out = ops.arma_filter(xd, coeff.to(cuda), dm.init_weight, dm.root_weight[0], dm.bias[0].reshape(K, F), plan,
                              x_root=rd)
        out.square().sum().backward()
        outs.append((out.detach().clone(), xd.grad.clone(), rd.grad.clone()))
    assert rel_err(outs[0][0], ref) <= TOL
    assert rel_err(outs[0][1], xo.grad) <= TOL
    assert rel_err(outs[0][2], ro.grad) <= TOL
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_arma_unsupported_options_raise(cuda):
    from feta_tmlr_b200 import ARMAConvDynamic
    with pytest.raises(NotImplementedError):
        ARMAConvDynamic(8, 8, num_stacks=2, num_layers=2)
    with pytest.raises(ValueError):
        ARMAConvDynamic(8, 16, num_stacks=2)
    mod = ARMAConvDynamic(8, 8, num_stacks=2).to(cuda)
    ei, batch, R = random_batch_graph(1, [4, 5])
    x = torch.randn(R, 8, device=cuda)
    with pytest.raises(ValueError):
        mod(x, ei.to(cuda), torch.randn(2, 3, device=cuda), batch=batch.to(cuda))
    with pytest.raises(NotImplementedError):
        mod(x, ei.to(cuda), torch.randn(2, 4, device=cuda), edge_weight=torch.ones(ei.shape[1], device=cuda),
            batch=batch.to(cuda))
