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


def test_lap_encoding_spans_reference_eigenvectors():
    gs = _graphs(8, shape='PATTERN')
    dim = 4
    out = pe.LapEncoding(dim, normalization='sym', device='cpu').compute_all(gs)
    for g, m in zip(gs, out):
        n = g['x'].shape[0]
        assert m.shape == (n, dim)
        L = od._dense_laplacian(torch.from_numpy(g['edge_index']), n, 'sym').astype(np.float64)
        w = np.sort(np.linalg.eigvalsh(L))
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


def test_synthetic_diffusion_pe_is_the_same_kernel():
    g = _graphs(1)[0]
    a = synthetic.diffusion_pe(g['edge_index'], g['x'].shape[0], 1.0)
    b = pe.DiffusionEncoding(None, 1.0, normalization='sym', device='cpu').compute_all([g])[0]
    assert np.allclose(a, b.numpy(), atol=2e-6)
