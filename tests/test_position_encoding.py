"""N3: batched position-encoding precompute vs the reference formulas (scipy expm / np.linalg.eig)."""
import numpy as np
import pytest
import torch

import oracle.data as od
from feta_tmlr_b200 import position_encoding as pe, synthetic


class _DS(list):
    pass


def _graphs(n=12, seed=3, shape='ZINC'):
    rng = np.random.default_rng(seed)
    return _DS(synthetic.make_graph(rng, shape) for _ in range(n))


@pytest.mark.parametrize("norm", [None, 'sym', 'rw'])
def test_diffusion_matches_scipy_expm(norm):
    gs = _graphs()
    out = pe.DiffusionEncoding(None, beta=1.0, normalization=norm, device='cpu').compute_all(gs)
    for g, m in zip(gs, out):
        ref = od.diffusion_pe(torch.from_numpy(g['edge_index']), g['x'].shape[0], 1.0, norm)
        assert m.dtype == torch.float32 and m.shape == ref.shape
        assert torch.allclose(m, ref.float(), atol=2e-6)


@pytest.mark.parametrize("norm,p", [('sym', 1), ('sym', 3), (None, 2), ('rw', 2)])
def test_pstep_matches_matrix_power(norm, p):
    gs = _graphs(6)
    out = pe.PStepRWEncoding(None, p=p, beta=0.5, normalization=norm, device='cpu').compute_all(gs)
    for g, m in zip(gs, out):
        ref = od.pstep_pe(torch.from_numpy(g['edge_index']), g['x'].shape[0], p, 0.5, norm)
        assert torch.allclose(m, ref.float(), atol=1e-5)


@pytest.mark.parametrize("norm", ['sym', None, 'rw'])
def test_lap_encoding_spans_reference_eigenvectors(norm):
    gs = _graphs(8, shape='PATTERN')
    dim = 4
    out = pe.LapEncoding(dim, normalization=norm, device='cpu').compute_all(gs)
    for g, m in zip(gs, out):
        n = g['x'].shape[0]
        assert m.shape == (n, dim)
        L = od._dense_laplacian(torch.from_numpy(g['edge_index']), n, norm).astype(np.float64)
        w = np.sort(np.linalg.eigvals(L).real)           # 'rw' is not symmetric: the reference's general eig
        # each column is a unit eigenvector of L for the matching (ascending, non-trivial) eigenvalue
        for c in range(dim):
            v = m[:, c].double().numpy()
            assert abs(np.linalg.norm(v) - 1.0) < 1e-5
            assert np.allclose(L @ v, w[c + 1] * v, atol=1e-4)


def test_cache_format_and_zero_diag(tmp_path):
    gs = _graphs(5)
    enc = pe.DiffusionEncoding(str(tmp_path / "zinc_diffusion_sym_1.0.pkl"), beta=1.0, normalization='sym',
                               zero_diag=True, device='cpu')
    enc.apply_to(gs, split='train')
    assert len(gs.pe_list) == 5 and float(gs.pe_list[0].diagonal().abs().max()) == 0.0
    import pickle
    with open(str(tmp_path / "zinc_diffusion_sym_1.0.pkl") + ".train", "rb") as f:
        cached = pickle.load(f)                           # position_encoding.py:35-49: list of dense tensors
    assert isinstance(cached, list) and float(cached[0].diagonal().abs().max()) > 0.0
    gs2 = _DS(gs)
    pe.DiffusionEncoding(str(tmp_path / "zinc_diffusion_sym_1.0.pkl"), beta=1.0, normalization='sym',
                         device='cpu').apply_to(gs2, split='train')
    assert torch.equal(gs2.pe_list[1], cached[1])
    full = pe.FullEncoding(None).compute_all(gs)
    assert full[0].shape == (gs[0]['x'].shape[0],) * 2
    assert pe.AdjEncoding(None).compute_all(gs)[0].shape[0] == 1


def test_rw_with_isolated_nodes_and_directed_edges():
    """'rw' rides on the 'sym' decomposition through L_rw = S^-1 L_sym S; isolated nodes (degree 0) are decoupled
    1x1 blocks, and a directed edge list must take the general per-graph formula."""
    g = dict(_graphs(1)[0])
    n = g['x'].shape[0]
    ei = g['edge_index']
    g['edge_index'] = ei[:, (ei[0] != 0) & (ei[1] != 0)]                 # node 0 isolated
    d = dict(g)
    d['edge_index'] = np.concatenate([g['edge_index'], np.array([[1], [n - 1]])], axis=1)   # + a one-way edge
    for gg in (g, d):
        e = torch.from_numpy(gg['edge_index'])
        m = pe.DiffusionEncoding(None, beta=1.0, normalization='rw', device='cpu').compute_all([gg])[0]
        assert torch.allclose(m, od.diffusion_pe(e, n, 1.0, 'rw').float(), atol=2e-6)
        m = pe.PStepRWEncoding(None, p=3, beta=0.5, normalization='rw', device='cpu').compute_all([gg])[0]
        assert torch.allclose(m, od.pstep_pe(e, n, 3, 0.5, 'rw').float(), atol=1e-5)
    got = pe.LapEncoding(4, normalization='rw', device='cpu').compute_all([d])[0]
    assert got.shape == (n, 4) and bool(torch.isfinite(got).all())


@pytest.mark.parametrize("norm", [None, 'sym', 'rw'])
def test_use_edge_attr_weights_the_laplacian(norm):
    """``use_edge_attr=True`` (position_encoding.py:66, :81, :129): ``graph.edge_attr`` is the edge weight of PyG
    ``get_laplacian``.  Symmetric positive weights (one per undirected edge) keep the batched path."""
    gs = _graphs(5)
    rng = np.random.default_rng(7)
    for g in gs:
        s, t = g['edge_index']
        key = np.minimum(s, t) * 1000 + np.maximum(s, t)
        w = {k: rng.uniform(0.5, 2.0) for k in np.unique(key)}
        g['edge_attr'] = np.array([w[k] for k in key])
    for g in gs:
        e, n, w = torch.from_numpy(g['edge_index']), g['x'].shape[0], torch.from_numpy(g['edge_attr']).float()
        m = pe.DiffusionEncoding(None, beta=1.0, use_edge_attr=True, normalization=norm, device='cpu').compute_all([g])[0]
        assert torch.allclose(m, od.diffusion_pe(e, n, 1.0, norm, edge_weight=w).float(), atol=5e-6)
        unweighted = pe.DiffusionEncoding(None, beta=1.0, normalization=norm, device='cpu').compute_all([g])[0]
        assert not torch.allclose(m, unweighted, atol=1e-3)
        m = pe.PStepRWEncoding(None, p=2, beta=0.5, use_edge_attr=True, normalization=norm, device='cpu').compute_all([g])[0]
        assert torch.allclose(m, od.pstep_pe(e, n, 2, 0.5, norm, edge_weight=w).float(), atol=1e-5)
        m = pe.LapEncoding(3, use_edge_attr=True, normalization=norm, device='cpu').compute_all([g])[0]
        L = od._dense_laplacian(e, n, norm, w).astype(np.float64)
        ev = np.sort(np.linalg.eigvals(L).real)
        for c in range(3):
            v = m[:, c].double().numpy()
            assert abs(np.linalg.norm(v) - 1.0) < 1e-5 and np.allclose(L @ v, ev[c + 1] * v, atol=1e-4)


def test_synthetic_diffusion_pe_is_the_same_kernel():
    g = _graphs(1)[0]
    a = synthetic.diffusion_pe(g['edge_index'], g['x'].shape[0], 1.0)
    b = pe.DiffusionEncoding(None, 1.0, normalization='sym', device='cpu').compute_all([g])[0]
    assert np.allclose(a, b.numpy(), atol=2e-6)


def test_directed_graph_takes_the_general_formula():
    """A non-symmetric Laplacian must not go through eigh (which reads one triangle)."""
    g = _graphs(1)[0]
    g = dict(g)
    g['edge_index'] = np.concatenate([g['edge_index'], np.array([[0], [g['x'].shape[0] - 1]])], axis=1)   # one-way edge
    out = pe.DiffusionEncoding(None, beta=1.0, normalization='sym', device='cpu').compute_all([g])[0]
    ref = od.diffusion_pe(torch.from_numpy(g['edge_index']), g['x'].shape[0], 1.0, 'sym')
    assert torch.allclose(out, ref.float(), atol=2e-6)
    # LapEncoding on a directed edge list: the reference's general np.linalg.eig, column by column up to sign
    got = pe.LapEncoding(4, normalization='sym', device='cpu').compute_all([g])[0]
    ref = od.lap_pe(torch.from_numpy(g['edge_index']), g['x'].shape[0], 4, 'sym')
    for c in range(4):
        assert min(float((got[:, c] - ref[:, c]).abs().max()), float((got[:, c] + ref[:, c]).abs().max())) < 1e-4


def test_chunked_batches_keep_dataset_order(monkeypatch):
    gs = _graphs(9, shape='PATTERN')
    whole = pe.DiffusionEncoding(None, beta=1.0, normalization='sym', device='cpu').compute_all(gs)
    monkeypatch.setattr(pe, 'CHUNK_BYTES', 3 * 190 * 190 * 8)           # forces several chunks
    parts = pe.DiffusionEncoding(None, beta=1.0, normalization='sym', device='cpu').compute_all(gs)
    for a, b, g in zip(whole, parts, gs):
        assert a.shape == (g['x'].shape[0],) * 2 and torch.allclose(a, b, atol=1e-6)


@pytest.mark.gpu
def test_position_encodings_on_cuda(cuda):
    """N3 on the device: batched eigh on the GPU vs the reference formulas (scipy expm / matrix power / eigvalsh)."""
    gs = _graphs(10)
    for norm in (None, 'sym'):
        out = pe.DiffusionEncoding(None, beta=1.0, normalization=norm, device='cuda').compute_all(gs)
        for g, m in zip(gs, out):
            ref = od.diffusion_pe(torch.from_numpy(g['edge_index']), g['x'].shape[0], 1.0, norm)
            assert m.device.type == 'cpu' and torch.allclose(m, ref.float(), atol=2e-6)
    out = pe.PStepRWEncoding(None, p=3, beta=0.5, normalization='sym', device='cuda').compute_all(gs)
    for g, m in zip(gs, out):
        ref = od.pstep_pe(torch.from_numpy(g['edge_index']), g['x'].shape[0], 3, 0.5, 'sym')
        assert torch.allclose(m, ref.float(), atol=1e-5)
    gp = _graphs(4, shape='CLUSTER')
    out = pe.LapEncoding(8, normalization='sym', device='cuda').compute_all(gp)
    for g, m in zip(gp, out):
        n = g['x'].shape[0]
        L = od._dense_laplacian(torch.from_numpy(g['edge_index']), n, 'sym').astype(np.float64)
        w = np.sort(np.linalg.eigvalsh(L))
        for c in range(8):
            v = m[:, c].double().numpy()
            assert abs(np.linalg.norm(v) - 1.0) < 1e-5 and np.allclose(L @ v, w[c + 1] * v, atol=1e-4)
