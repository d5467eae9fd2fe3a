"""GPU parity: coefficient path, pack/unpack, pooling and the GPU batch builder vs the oracle."""
import numpy as np
import pytest
import torch

from helpers import make_batch, rel_err, to_dev
from oracle import pyg17
from oracle.dense import dense_coeff_scalar
from oracle.models import OracleEncoderGenGCN, OracleGlobalAvg1D
from oracle.layers import OracleDiffTransformerEncoderLayer

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _attn(seed, B, H, nmax, lens, zero_frac=0.3):
    g = torch.Generator().manual_seed(seed)
    lens = torch.as_tensor(lens)
    mask = torch.arange(nmax)[None, :] >= lens[:, None]
    a = torch.rand(B, H, nmax, nmax, generator=g)
    a = a * (torch.rand(B, H, nmax, nmax, generator=g) > zero_frac)       # exact zeros incl. some diagonals
    valid = (~mask)[:, None, :, None] & (~mask)[:, None, None, :]
    a = a * valid
    a = a / a.sum(-1, keepdim=True).clamp(min=1e-6)
    return a, mask


@pytest.mark.parametrize("lens", [[5, 3, 7, 1], [38, 9, 23], [188, 44]])
def test_coeff_scalar_matches_dense_closed_form(cuda, lens):
    from feta_tmlr_b200 import ops
    B, H, nmax = len(lens), 4, max(lens)
    a, mask = _attn(1, B, H, nmax, lens)
    node_ptr = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
    N = int(sum(lens))
    s = ops.coeff_scalar(a.to(cuda), mask.to(cuda), node_ptr.to(cuda), N).cpu()
    for h in range(H):
        for b in range(B):
            n = lens[b]
            ref = dense_coeff_scalar(a[b, h, :n, :n].double()).float()
            got = s[h * N + int(node_ptr[b]): h * N + int(node_ptr[b]) + n]
            assert rel_err(got, ref) < TOL


def test_filter_coefficients_match_literal_all_pairs_gcn(cuda):
    """models.py:240-287 restated literally (host loop + all-pairs GCNConv) vs the collapsed kernels,
    values and gradients of gcn.weight / gcn.bias / linear.*"""
    from feta_tmlr_b200 import DiffTransformerEncoderGenGCN, DiffTransformerEncoderLayer
    torch.manual_seed(0)
    d, H, lens = 16, 2, [5, 3, 7, 1]
    B, nmax = len(lens), max(lens)
    o = OracleEncoderGenGCN(d, H, OracleDiffTransformerEncoderLayer(d, H, 2 * d, 0.0), 1)
    m = DiffTransformerEncoderGenGCN(d, H, DiffTransformerEncoderLayer(d, H, 2 * d, 0.0), 1).to(cuda)
    m.load_state_dict(o.state_dict())
    a, mask = _attn(2, B, H, nmax, lens)
    fi = torch.tensor([[b, i] for b in range(B) for i in range(lens[b])])
    batch = fi[:, 0].clone()
    ei = torch.zeros((2, 0), dtype=torch.long)
    co = o.get_filter_coefficients(a, None, None, None, mask)
    cg = m.get_filter_coefficients(a.to(cuda), ei.to(cuda), fi.to(cuda), batch.to(cuda), mask.to(cuda))
    assert co.shape == cg.shape and rel_err(cg, co) < TOL
    w = torch.randn(co.shape, generator=torch.Generator().manual_seed(1))
    (co * w).sum().backward()
    (cg * w.to(cuda)).sum().backward()
    for name in ["gcn.weight", "gcn.bias", "linear.weight", "linear.bias"]:
        po, pg = dict(o.named_parameters())[name], dict(m.named_parameters())[name]
        assert rel_err(pg.grad, po.grad) < 2e-4, name


def test_pack_unpack_heads(cuda):
    from feta_tmlr_b200 import ops
    B, nmax, H, dh, lens = 3, 6, 4, 8, [6, 2, 4]
    fi = torch.tensor([[b, i] for b in range(B) for i in range(lens[b])])
    N = fi.shape[0]
    oh = torch.randn(B, nmax, H, dh)
    # reference: models.py:179-185 + :347
    out_heads = oh.permute([2, 0, 1, 3]).reshape(H * B, nmax, dh)
    fia = fi.repeat(H, 1)
    fia[:, 0] += torch.arange(H).repeat_interleave(N) * B
    ref = out_heads[fia[:, 0], fia[:, 1], :]
    ohg = oh.to(cuda).requires_grad_()
    x = ops.pack_heads(ohg, fi.to(cuda))
    assert torch.equal(x.cpu(), ref)                                      # pure data movement: bit exact
    w = torch.randn(H * N, dh)
    x.backward(w.to(cuda))
    oh2 = oh.clone().requires_grad_()
    oh2.permute([2, 0, 1, 3]).reshape(H * B, nmax, dh)[fia[:, 0], fia[:, 1], :].backward(w)
    assert torch.equal(ohg.grad.cpu(), oh2.grad)
    # reference: models.py:200-202
    y = torch.randn(H * N, dh)
    filt = y.reshape(H, N, dh).permute(1, 0, 2).reshape(N, H * dh)
    ref2 = torch.zeros(nmax, B, H * dh)
    ref2[fi[:, 1], fi[:, 0], :] = filt
    yg = y.to(cuda).requires_grad_()
    out = ops.unpack_heads(yg, fi.to(cuda), B, nmax, H)
    assert torch.equal(out.cpu(), ref2)
    w2 = torch.randn(nmax, B, H * dh)
    out.backward(w2.to(cuda))
    refg = w2[fi[:, 1], fi[:, 0], :].reshape(N, H, dh).permute(1, 0, 2).reshape(H * N, dh)
    assert torch.equal(yg.grad.cpu(), refg)


def test_pooling_ops(cuda):
    from feta_tmlr_b200 import ops, GlobalAvg1D
    sizes = [3, 1, 7, 40]
    gp = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    x = torch.randn(sum(sizes), 33)
    xo, xg = x.clone().requires_grad_(), x.to(cuda).requires_grad_()
    ro = pyg17.global_mean_pool(xo, batch)
    rg = ops.segment_mean(xg, gp.to(cuda))
    assert rel_err(rg, ro) < 1e-6
    ro.square().sum().backward()
    rg.square().sum().backward()
    assert rel_err(xg.grad, xo.grad) < 1e-6
    B, nmax, C = 4, 9, 20
    lens = torch.tensor([9, 1, 5, 3])
    mask = torch.arange(nmax)[None, :] >= lens[:, None]
    xs = torch.randn(nmax, B, C)                                          # seq-first storage, permuted view
    xo, xg = xs.clone().requires_grad_(), xs.to(cuda).requires_grad_()
    ro = OracleGlobalAvg1D()(xo.permute(1, 0, 2), mask)
    rg = GlobalAvg1D()(xg.permute(1, 0, 2), mask.to(cuda))
    assert rel_err(rg, ro) < 1e-6
    ro.square().sum().backward()
    rg.square().sum().backward()
    assert rel_err(xg.grad, xo.grad) < 1e-6
    fi = torch.tensor([[b, i] for b in range(B) for i in range(int(lens[b]))])
    po = xs.permute(1, 0, 2)[~mask]                                       # models.py:1070-1071
    pg = ops.gather_rows(xs.to(cuda).permute(1, 0, 2), fi.to(cuda))
    assert torch.equal(pg.cpu(), po)


@pytest.mark.parametrize("name", ["MUTAG", "ZINC", "PATTERN", "CLUSTER", "MOLHIV"])
def test_device_batch_builder_bit_exact(cuda, name):
    """A7: GPU batch builder == host collate == reference collate loops, bit for bit."""
    from feta_tmlr_b200 import data as fdata
    ids = np.array([7, 2, 11, 0, 5, 9])
    cfg, graphs, store, host = make_batch(name, 12, seed=3, ids=ids)
    dev = fdata.DeviceBatchBuilder(store, cuda).build(ids)
    torch.cuda.synchronize()
    assert len(dev) >= 9
    for a, b in zip(dev, host):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.dtype == b.dtype and torch.equal(a.cpu(), b), name


@pytest.mark.parametrize("name", ["ZINC", "PATTERN", "CLUSTER", "MOLHIV"])
def test_device_batch_builder_static_mode(cuda, name):
    """N2 for the CUDA-graph engine: ``build(ids, static=caps)`` == ``collate_host(..., static=caps)`` bit for bit,
    and a batch that does not fit the capacities raises."""
    from feta_tmlr_b200 import data as fdata, engine
    ids = np.array([7, 2, 11, 0, 5, 9])
    cfg, graphs, store, _ = make_batch(name, 12, seed=3, ids=ids)
    caps = engine.static_caps(store, len(ids))
    caps = (caps[0] + 2, caps[1])
    host = fdata.collate_host(store, ids, static=caps)
    dev = fdata.DeviceBatchBuilder(store, cuda).build(ids, static=caps)
    torch.cuda.synchronize()
    for k, (a, b) in enumerate(zip(dev, host)):
        assert (a is None) == (b is None), k
        if a is not None:
            assert a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape) and torch.equal(a.cpu(), b), (name, k)
    with pytest.raises(ValueError):
        fdata.DeviceBatchBuilder(store, cuda).build(ids, static=(caps[0], 8))


def test_graphed_step_fed_by_device_batch_builder(cuda):
    """The GPU batch builder feeds the whole-step CUDA graph: same losses as host-collated static batches."""
    import copy
    import feta_tmlr_b200.models as fmodels
    from feta_tmlr_b200 import data as fdata, engine, synthetic
    name, B = "ZINC", 8
    cfg = synthetic.CONFIGS[name]
    graphs = synthetic.make_dataset(name, 3 * B, seed=9)
    store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
    caps = engine.static_caps(store, B)
    idsets = [np.arange(i * B, (i + 1) * B) for i in range(3)]
    host = [fdata.collate_host(store, ids, static=caps) for ids in idsets]
    torch.manual_seed(0)
    m1 = synthetic.build_model(name, fmodels, layers=2).to(cuda)
    m2 = copy.deepcopy(m1)
    lf = torch.nn.functional.l1_loss
    e1 = engine.GraphedTrainStep(m1, lf, host[0], lr=1e-3, device=cuda)
    builder = fdata.DeviceBatchBuilder(store, cuda)
    e2 = engine.GraphedTrainStep(m2, lf, builder.build(idsets[0], static=caps), lr=1e-3, device=cuda)
    for i in range(3):
        l1 = float(e1.step(host[i]))
        l2 = float(e2.step(builder.build(idsets[i], static=caps)))
        assert abs(l1 - l2) <= 1e-6 * max(1.0, abs(l1)), (i, l1, l2)


@pytest.mark.parametrize("edge_dtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("tile", [False, True])
@pytest.mark.parametrize("B,nmax,H", [(5, 9, 2), (128, 37, 8), (1000, 40, 4)])
def test_static_context_tensors_bit_exact(cuda, B, nmax, H, tile, edge_dtype):
    """feta_static_context (two launches) == the tensor-op formulation of the padded-domain context: graph sizes,
    packed id -> padded slot id of every edge endpoint (searchsorted over the size prefix), (-1, -1) padding columns,
    pooling segments, real-row weights."""
    from feta_tmlr_b200 import ops
    g = torch.Generator().manual_seed(B + nmax)
    lens = torch.randint(1, nmax + 1, (B,), generator=g)
    lens[0] = nmax
    masks = torch.arange(nmax)[None, :] >= lens[:, None]
    node_ptr = torch.cumsum(lens, 0)
    first = node_ptr - lens
    cols = []
    for b in range(B):
        e = int(torch.randint(0, 3 * int(lens[b]) + 1, (1,), generator=g))
        cols.append(first[b] + torch.randint(0, int(lens[b]), (2, e), generator=g))
    edges = torch.cat(cols, dim=1)
    ecap = edges.shape[1] + 17
    ei_in = torch.full((2, ecap), -1, dtype=torch.int64)
    ei_in[:, :edges.shape[1]] = edges
    # tensor-op formulation (what forward_static ran before)
    pad = ei_in < 0
    b_of = torch.searchsorted(node_ptr, ei_in, right=True).clamp_(max=B - 1)
    ref = ei_in - first[b_of] + b_of * nmax
    if tile:
        heads = torch.arange(H, dtype=torch.int64)
        ref = (ref.view(2, 1, -1) + (heads * B * nmax).view(1, H, 1)).reshape(2, -1)
        pad = pad.view(2, 1, -1).expand(2, H, -1).reshape(2, -1)
    ref = ref.masked_fill(pad, -1)
    gidx = torch.arange(H * B)
    ei, slot_ptr, seg_lo, seg_hi, real = ops.static_context_tensors(masks.to(cuda), ei_in.to(edge_dtype).to(cuda), nmax,
                                                                     H, tile_heads=tile)
    assert ei.dtype == torch.int64 and torch.equal(ei.cpu(), ref)
    assert torch.equal(slot_ptr.cpu().long(), torch.arange(B) * nmax)
    assert torch.equal(seg_lo.cpu().long(), gidx * nmax)
    assert torch.equal(seg_hi.cpu().long(), gidx * nmax + lens.repeat(H))
    assert torch.equal(real.cpu(), (~masks).t().unsqueeze(-1).float())
