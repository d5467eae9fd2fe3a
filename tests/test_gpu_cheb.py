"""GPU parity: CSR plan + fused ChebConvDynamic vs the CPU oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from helpers import oracle_csr, random_batch_graph, rel_err
from oracle.cheb import OracleChebConvDynamic, cheb_conv_dynamic

pytestmark = pytest.mark.gpu
TOL = 1e-4      # north_star: <= 1e-4 relative, fp32


def _plan(cuda, ei, batch, R, G, hints=None):
    from feta_tmlr_b200 import ops
    return ops.build_cheb_plan(ei.to(cuda), None if batch is None else batch.to(cuda), R, G, 2.0, hints=hints)


@pytest.mark.parametrize("sizes", [[5, 7, 1, 4], [1], [33, 2, 64, 17, 1, 1, 90], [200, 150]])
@pytest.mark.parametrize("bdtype", [torch.int64, torch.float32])
def test_plan_bit_exact(cuda, sizes, bdtype):
    ei, batch, R = random_batch_graph(1, sizes)
    p = _plan(cuda, ei, batch.to(bdtype), R, len(sizes))
    m = p.meta_host()
    for tr, (rp, ci, va) in ((False, (p.rowptr, p.colidx, p.vals)), (True, (p.rowptr_t, p.colidx_t, p.vals_t))):
        orp, oci, ova = oracle_csr(ei, R, transpose=tr)
        nnz = int(orp[-1])
        assert m[0] == nnz
        assert np.array_equal(rp.cpu().numpy(), orp)                         # bit-exact integer work
        assert np.array_equal(ci.cpu().numpy()[:nnz], oci)
        np.testing.assert_allclose(va.cpu().numpy()[:nnz], ova, rtol=2e-7, atol=0)
    gp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    assert np.array_equal(p.graph_ptr.cpu().numpy(), gp)
    assert np.array_equal(p.row_graph.cpu().numpy()[:R], np.repeat(np.arange(len(sizes)), sizes))
    assert m[1] == len(sizes) and m[2] == 1 and m[3] == 1 and m[4] == max(sizes)


def test_plan_flags(cuda):
    from feta_tmlr_b200 import ops
    ei, batch, R = random_batch_graph(2, [4, 5])
    with pytest.raises(ValueError):                                   # unsorted batch
        _plan(cuda, ei, torch.tensor([1, 0, 0, 0, 1, 1, 1, 1, 1]), R, 2)
    with pytest.raises(RuntimeError):                                 # graph-count mismatch
        _plan(cuda, ei, batch, R, 3)
    bad = torch.tensor([[0, 99], [1, 2]])
    with pytest.raises(IndexError):
        _plan(cuda, bad, batch, R, 2)
    cross = torch.tensor([[0, 5], [5, 0]])                            # edge between the two graphs
    p = _plan(cuda, cross, batch, R, 2)
    assert p.block_diagonal is False
    e = _plan(cuda, torch.zeros((2, 0), dtype=torch.int64), batch, R, 2)   # no edges at all
    assert e.meta_host()[0] == 0


def _run_both(cuda, sizes, F, K, seed, float_batch=True, directed_extra=0, hints=None, cross=False):
    from feta_tmlr_b200 import ChebConvDynamic
    ei, batch, R = random_batch_graph(seed, sizes, directed_extra=directed_extra)
    if cross:
        ei = torch.cat([ei, torch.tensor([[0, R - 1], [R - 1, 0]])], dim=1)
    G = len(sizes)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, F, generator=g)
    coeff = torch.randn(G, K * F * F, generator=g) * 0.3              # contiguous [G, K*F*F] (models.py:357)
    bias = torch.randn(F, generator=g)
    dout = torch.randn(R, F, generator=g)
    b = batch.float() if float_batch else batch

    xo, co = x.clone().requires_grad_(), coeff.clone().requires_grad_()
    om = OracleChebConvDynamic(F, F, K)
    om.bias.data.copy_(bias)
    oo = om(xo, ei, co.reshape(-1, K, F, F).permute(1, 0, 2, 3), batch=b)
    oo.reshape(R, F).backward(dout)

    m = ChebConvDynamic(F, F, K).to(cuda)
    m.bias.data.copy_(bias)
    m.plan_hints = hints
    xg, cg = x.to(cuda).requires_grad_(), coeff.to(cuda).requires_grad_()
    og = m(xg, ei.to(cuda), cg.reshape(-1, K, F, F).permute(1, 0, 2, 3), batch=b.to(cuda))
    og.reshape(R, F).backward(dout.to(cuda))
    torch.cuda.synchronize()
    assert og.shape == oo.shape
    return (rel_err(og, oo), rel_err(xg.grad, xo.grad), rel_err(cg.grad, co.grad),
            rel_err(m.bias.grad, om.bias.grad)), m


@pytest.mark.parametrize("F", [4, 8, 16, 32])
@pytest.mark.parametrize("K", [1, 2, 4])
def test_cheb_fused_parity(cuda, F, K):
    errs, _ = _run_both(cuda, [5, 7, 1, 4, 38, 23, 1, 2, 64], F, K, seed=F * 10 + K)
    assert max(errs) < TOL, errs


@pytest.mark.parametrize("sizes", [[1], [2], [188, 44, 120, 97, 188], [300, 10], [1] * 70 + [3] * 50])
def test_cheb_graph_shapes(cuda, sizes):
    errs, _ = _run_both(cuda, sizes, 16, 4, seed=len(sizes))
    assert max(errs) < TOL, errs


def test_cheb_directed_edges_use_transposed_csr(cuda):
    errs, _ = _run_both(cuda, [9, 12, 30], 8, 4, seed=5, directed_extra=6)
    assert max(errs) < TOL, errs


def test_cheb_int_batch_and_hints(cuda):
    errs, m = _run_both(cuda, [9, 12, 30], 16, 3, seed=6, float_batch=False,
                        hints={'max_nodes': 64, 'block_diagonal': True})
    assert max(errs) < TOL, errs
    plan = next(iter(m._plans.values()))
    assert plan.validate()[7] == 0


@pytest.mark.parametrize("case", ["odd_F", "huge_graph", "cross_edges"])
def test_cheb_unfused_fallback(cuda, case):
    if case == "odd_F":
        errs, _ = _run_both(cuda, [5, 7, 11], 6, 4, seed=7)
    elif case == "huge_graph":
        errs, _ = _run_both(cuda, [1500, 20], 16, 3, seed=8)
    else:
        errs, _ = _run_both(cuda, [5, 7, 11], 8, 4, seed=9, cross=True)
    assert max(errs) < TOL, errs


def test_cheb_guard_refuses_bad_hints(cuda):
    """Hints that under-state the largest graph must not fault the GPU: the kernel refuses."""
    from feta_tmlr_b200 import ChebConvDynamic
    ei, batch, R = random_batch_graph(3, [100, 20])
    m = ChebConvDynamic(8, 8, 2).to(cuda)
    m.plan_hints = {'max_nodes': 16, 'block_diagonal': True}
    theta = torch.randn(2, 2, 8, 8, device=cuda)
    m(torch.randn(R, 8, device=cuda), ei.to(cuda), theta, batch=batch.to(cuda))
    plan = next(iter(m._plans.values()))
    with pytest.raises(RuntimeError):
        plan.validate()


def test_cheb_known_answers(cuda):
    """Closed forms: isolated rows give (x, 0, -x, 0); P2 alternates."""
    from feta_tmlr_b200 import ChebConvDynamic
    F, K = 4, 4
    eye = torch.eye(F)
    m = ChebConvDynamic(F, F, K, bias=False).to(cuda)
    x = torch.randn(6, F)
    batch = torch.tensor([0, 0, 1, 1, 1, 1])
    ei = torch.tensor([[0, 1], [1, 0]])                                 # P2 on graph 0; graph 1 isolated
    for k in range(K):
        theta = torch.zeros(K, 2, F, F)
        theta[k] = eye
        out = m(x.to(cuda), ei.to(cuda), theta.to(cuda), batch=batch.to(cuda)).cpu()
        iso = [x[2:], torch.zeros(4, F), -x[2:], torch.zeros(4, F)][k]
        assert torch.allclose(out[2:], iso, atol=1e-6)
        p2 = x[:2] if k % 2 == 0 else -x[:2].flip(0)                     # L_hat = [[0,-1],[-1,0]]
        assert torch.allclose(out[:2], p2, atol=1e-6)


def test_cheb_learn_only_mode(cuda):
    from feta_tmlr_b200 import ChebConvDynamic
    ei, batch, R = random_batch_graph(4, [6, 9, 3])
    F, K, G = 8, 3, 3
    om = OracleChebConvDynamic(F, F, K, learn_only_filter_order_coeff=True)
    m = ChebConvDynamic(F, F, K, learn_only_filter_order_coeff=True).to(cuda)
    m.load_state_dict(om.state_dict())
    x = torch.randn(R, F)
    c = torch.randn(K, G)
    xo, co = x.clone().requires_grad_(), c.clone().requires_grad_()
    oo = om(xo, ei, co, batch=batch.float())
    oo.sum().backward()
    xg, cg = x.to(cuda).requires_grad_(), c.to(cuda).requires_grad_()
    og = m(xg, ei.to(cuda), cg, batch=batch.float().to(cuda))
    og.sum().backward()
    assert rel_err(og, oo) < TOL and rel_err(xg.grad, xo.grad) < TOL and rel_err(cg.grad, co.grad) < TOL
    assert rel_err(m.weight.grad, om.weight.grad) < TOL


def test_cheb_rejects_unsupported(cuda):
    from feta_tmlr_b200 import ChebConvDynamic
    m = ChebConvDynamic(4, 4, 2, normalization='rw').to(cuda)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(2, 4, device=cuda), torch.zeros(2, 0, dtype=torch.long, device=cuda),
          torch.zeros(2, 1, 4, 4, device=cuda), batch=torch.zeros(2, device=cuda), lambda_max=2.0)
    m = ChebConvDynamic(4, 4, 2).to(cuda)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 4), torch.zeros(2, 0, dtype=torch.long), torch.zeros(2, 1, 4, 4), batch=torch.zeros(2))
    m = ChebConvDynamic(4, 4, 2).to(cuda)
    with pytest.raises(NotImplementedError):                          # __norm__ keeps a diagonal unless lambda_max = 2
        m(torch.zeros(2, 4, device=cuda), torch.zeros(2, 0, dtype=torch.long, device=cuda),
          torch.zeros(2, 1, 4, 4, device=cuda), batch=torch.zeros(2, device=cuda), lambda_max=3.0)


@pytest.mark.parametrize("sizes", [[5, 9, 7], [80, 120]])
def test_guard_refusal_is_visible(cuda, sizes):
    """Hinted plans skip the host-side validation; when the hints do not hold the fused kernels refuse to run,
    and they must leave NaN (never uninitialised memory) in every output, forward and backward."""
    from feta_tmlr_b200 import ops
    ei, batch, R = random_batch_graph(7, sizes)
    G, F, K = len(sizes), 16, 4
    plan = ops.build_cheb_plan(ei.to(cuda), batch.to(cuda), R, G, 2.0,
                               hints={'max_nodes': min(sizes) - 1, 'block_diagonal': True})   # wrong on purpose
    x = torch.randn(R, F, device=cuda, requires_grad=True)
    theta = torch.randn(K, G, F, F, device=cuda, requires_grad=True)
    bias = torch.zeros(F, device=cuda)
    out = ops.cheb_filter(x, theta, bias, plan)
    assert bool(torch.isnan(out).all())
    out.backward(torch.ones_like(out))
    assert bool(torch.isnan(x.grad).all()) and bool(torch.isnan(theta.grad).all())
    with pytest.raises(RuntimeError):
        plan.validate()


@pytest.mark.parametrize("family", ["tile", "no_dense", "no_tile_no_dense", "warp_v1", "chunk_only"])
@pytest.mark.parametrize("F,K,sizes", [(16, 4, [5, 7, 1, 4, 38, 23, 1, 2, 64]), (8, 4, [30, 31, 32, 33, 9, 12] * 6),
                                       (16, 3, [90, 17, 128, 1, 66]), (8, 2, [200, 129, 256, 70]),
                                       (16, 4, [188, 44, 120, 97, 188])])
def test_cheb_every_kernel_family(cuda, monkeypatch, family, F, K, sizes):
    """The dispatcher picks one of five forward families (tcgen05 tile kernel, CTA-per-graph dense kernel, the two
    generations of the warp-per-graph kernel -- cheb_lane.cu, cheb_warp.cu --, chunk kernel) by shape; every family is
    forced here on shapes the others would take."""
    env = {"tile": {"FETA_CHEB_TILE": "1", "FETA_CHEB_NO_DENSE_KERNEL": "1"},
           "no_dense": {"FETA_CHEB_NO_DENSE_KERNEL": "1"},
           "no_tile_no_dense": {"FETA_CHEB_NO_DENSE_KERNEL": "1", "FETA_CHEB_NO_TILE_KERNEL": "1"},
           "warp_v1": {"FETA_CHEB_NO_DENSE_KERNEL": "1", "FETA_CHEB_NO_TILE_KERNEL": "1",
                       "FETA_CHEB_NO_LANE_KERNEL": "1"},
           "chunk_only": {"FETA_CHEB_NO_DENSE_KERNEL": "1", "FETA_CHEB_NO_TILE_KERNEL": "1",
                          "FETA_CHEB_NO_WARP_KERNEL": "1"}}[family]
    for k in ("FETA_CHEB_TILE", "FETA_CHEB_NO_DENSE_KERNEL", "FETA_CHEB_NO_TILE_KERNEL", "FETA_CHEB_NO_WARP_KERNEL",
              "FETA_CHEB_NO_LANE_KERNEL"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    errs, _ = _run_both(cuda, sizes, F, K, seed=F + K + len(sizes))
    assert max(errs) < TOL, (family, errs)
