"""Golden vectors produced by RUNNING THE UNMODIFIED REFERENCE (``/root/reference/transformer/{ChebNetDynamic,
models,data}.py``) under ``ref_shim`` -- the pin of the oracle and of the CUDA path (tier rule 3).

    python tests/golden/make_golden_from_reference.py          # needs /root/reference; writes tests/golden/ref_*.pt.gz

What is reference code and what is stand-in is listed at the top of ``ref_shim.py``: the Chebyshev operator, the
encoder (head stacking, ``get_filter_coefficients`` with its host loop and literal all-pairs ``GCNConv``, ``filter``,
scatter-back, ``linear_cat``), the three model heads, ``GlobalAvg1D`` and the three collates are executed from the
reference's own files; PyG-1.7 primitives come from the shim; the attention layer -- absent from the reference
tree -- is ``oracle/layers.py``.

Floating-point cases run the reference in float64 (``torch.set_default_dtype``: the reference allocates several
temporaries with the default dtype, e.g. models.py:201,280) on float32-representable inputs, so the stored
outputs are accurate to ~1e-15 and the fp32 CUDA path is judged against them at 1e-4.  Weights are never stored:
``helpers.det_init`` regenerates them from a seed.  Integer cases (collate, ``__norm__`` edge lists) are stored
verbatim and compared bit-exactly.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from helpers import BIG_GRAD, det_init, grad_summary, random_batch_graph, save_fixture  # noqa: E402
from feta_tmlr_b200 import synthetic  # noqa: E402

REF = ref_shim.install()


class f64(object):
    def __enter__(self):
        torch.set_default_dtype(torch.float64)

    def __exit__(self, *a):
        torch.set_default_dtype(torch.float32)


def _randn(rs, *shape, scale=1.0):
    return torch.from_numpy((rs.standard_normal(shape) * scale).astype(np.float32))


# ---------------------------------------------------------------------------------------------------
# A1/A2/A3: ChebConvDynamic.forward / __norm__ / message   (ChebNetDynamic.py:108-193)
# ---------------------------------------------------------------------------------------------------
def cheb_cases():
    specs = [
        # name, sizes, F, K, heads(H: rows of heads 1.. are isolated), float batch, learn_only, extras
        dict(name="mol_f8_k4", sizes=[9, 23, 1, 14, 37, 12], F=8, K=4, H=1, fbatch=True),
        dict(name="mol_f16_k4_heads", sizes=[17, 28, 10, 21], F=16, K=4, H=4, fbatch=True),
        dict(name="sbm_f16_k4", sizes=[70, 131], F=16, K=4, H=2, fbatch=True, deg=20.0),
        dict(name="k1", sizes=[5, 6], F=8, K=1, H=1, fbatch=False),
        dict(name="k2", sizes=[5, 1, 6], F=4, K=2, H=1, fbatch=False),
        dict(name="k3_directed", sizes=[8, 11, 3], F=32, K=3, H=1, fbatch=False, directed=3),
        dict(name="learn_only_k4", sizes=[12, 7, 19], F=16, K=4, H=2, fbatch=True, learn_only=True),
    ]
    out = []
    for i, s in enumerate(specs):
        rs = np.random.RandomState(100 + i)
        ei, batch, N = random_batch_graph(100 + i, s['sizes'], deg=s.get('deg', 3.0),
                                          directed_extra=s.get('directed', 0))
        H, Fc, K, B = s['H'], s['F'], s['K'], len(s['sizes'])
        R, G = H * N, H * B
        batch_all = torch.cat([batch + h * B for h in range(H)])          # models.py:181-182 (edges stay un-tiled)
        x = _randn(rs, R, Fc)
        w = _randn(rs, R, Fc)
        bias = _randn(rs, Fc, scale=0.3)
        learn_only = s.get('learn_only', False)
        with f64():
            m = REF.cheb.ChebConvDynamic(Fc, Fc, K, learn_only_filter_order_coeff=learn_only)
            m.bias.data.copy_(bias.double())
            xd = x.double().requires_grad_()
            if learn_only:
                weight = _randn(rs, K, Fc, Fc, scale=0.3)
                m.weight.data.copy_(weight.double())
                coeff = _randn(rs, G, K)
                cd = coeff.double().requires_grad_()
                fc = cd.reshape((-1, K)).permute([1, 0])                                 # models.py:359
            else:
                weight = None
                coeff = _randn(rs, G, K * Fc * Fc, scale=0.3)
                cd = coeff.double().requires_grad_()
                fc = cd.reshape((-1, K, Fc, Fc)).permute([1, 0, 2, 3])                   # models.py:357
            b = batch_all.double() if s['fbatch'] else batch_all
            y = m(xd, ei, fc, batch=b)
            (y * w.double()).sum().backward()
            n_ei, n_w = m.__norm__(ei, R, None, 'sym', torch.tensor(2.0), dtype=torch.float64, batch=b)
        out.append(dict(name=s['name'], F=Fc, K=K, H=H, B=B, N=N, learn_only=learn_only, float_batch=s['fbatch'],
                        edge_index=ei, batch=batch_all, x=x, w=w, bias=bias, coeff=coeff, weight=weight,
                        out=y.detach(), dx=xd.grad, dcoeff=cd.grad, dbias=m.bias.grad.clone(),
                        dweight=None if not learn_only else m.weight.grad.clone(),
                        norm_edge_index=n_ei, norm_weight=n_w))
    return out


# ---------------------------------------------------------------------------------------------------
# N4: ARMAConvDynamic.forward   (ChebNetDynamic.py:297-346)
# ---------------------------------------------------------------------------------------------------
def arma_cases():
    out = []
    for i, (sizes, Fc, K, H) in enumerate([([9, 14, 1, 22], 8, 4, 2), ([30, 75], 16, 4, 1)]):
        rs = np.random.RandomState(200 + i)
        ei, batch, N = random_batch_graph(200 + i, sizes, deg=3.0 if i == 0 else 12.0, self_loops=False)
        B = len(sizes)
        R, G = H * N, H * B
        batch_all = torch.cat([batch + h * B for h in range(H)])
        x, w = _randn(rs, R, Fc), _randn(rs, R, Fc)
        coeff = _randn(rs, G, 2 * K)
        with f64():
            m = REF.cheb.ARMAConvDynamic(Fc, Fc, num_stacks=K, num_layers=1)
            det_init(m, 200 + i)
            xd, cd = x.double().requires_grad_(), coeff.double().requires_grad_()
            y = m(xd, ei, cd, batch=batch_all.double())
            (y * w.double()).sum().backward()
        out.append(dict(F=Fc, K=K, H=H, B=B, N=N, seed=200 + i, edge_index=ei, batch=batch_all, x=x, w=w, coeff=coeff,
                        out=y.detach(), dx=xd.grad, dcoeff=cd.grad,
                        grads={k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    return out


# ---------------------------------------------------------------------------------------------------
# A4: get_filter_coefficients (models.py:240-287)  +  A5: GlobalAvg1D (models.py:586-595)
# ---------------------------------------------------------------------------------------------------
def coeff_cases():
    out = []
    for i, (lens, H, dh) in enumerate([([7, 3, 9, 1], 2, 4), ([25, 12, 33], 4, 8)]):
        rs = np.random.RandomState(300 + i)
        B, nmax, d = len(lens), max(lens), H * dh
        lens_t = torch.tensor(lens)
        mask = torch.arange(nmax)[None, :] >= lens_t[:, None]
        attn = torch.from_numpy(rs.uniform(0, 1, size=(B, H, nmax, nmax)).astype(np.float32))
        attn[torch.from_numpy(rs.uniform(size=attn.shape) < 0.15)] = 0.0        # exact zeros are dropped (:276)
        attn = attn / attn.sum(-1, keepdim=True).clamp_min(1e-6)
        wout = None
        with f64():
            layer = REF.models.DiffTransformerEncoderLayer(d, H, 2 * d, 0.0)
            enc = REF.models.DiffTransformerEncoderGenGCN(d, H, layer, 1, num_coefficients=4)
            det_init(enc, 300 + i)
            c = enc.get_filter_coefficients(attn.double(), None, None, None, mask)      # [H, B, ncoef]
            wout = _randn(rs, *c.shape)
            (c * wout.double()).sum().backward()
            grads = {k: (grad_summary(p.grad) if (p.grad.numel() > BIG_GRAD and p.grad.dim() == 2) else p.grad.clone())
                     for k, p in enc.named_parameters() if p.grad is not None}
        out.append(dict(H=H, dh=dh, lens=lens, seed=300 + i, attn=attn, mask=mask, w=wout, coeff=c.detach(),
                        grads=grads))
    rs = np.random.RandomState(310)
    x = _randn(rs, 5, 11, 12)
    mask = torch.arange(11)[None, :] >= torch.tensor([11, 3, 7, 1, 9])[:, None]
    with f64():
        pooled = REF.models.GlobalAvg1D()(x.double(), mask)
    return out, dict(x=x, mask=mask, out=pooled)


# ---------------------------------------------------------------------------------------------------
# A7: the three collates (data.py:161-225, :277-344, :394-460)
# ---------------------------------------------------------------------------------------------------
def _ref_dataset(name, graphs):
    cfg = synthetic.CONFIGS[name]
    with_edge_attr = cfg['kind'] == 'ogb'          # data.py:342 concatenates g.edge_attr unconditionally
    rs = np.random.RandomState(7)
    datas = []
    for g in graphs:
        x = torch.from_numpy(np.asarray(g['x']))
        y = torch.as_tensor(g['y'])
        ea = None
        if with_edge_attr:
            ea = torch.from_numpy(rs.randint(0, 2, size=(g['edge_index'].shape[1], 3)).astype(np.int64))
        datas.append(REF.Data(x, torch.from_numpy(np.asarray(g['edge_index'], dtype=np.int64)), y, ea))
    cls = {'v2': REF.data.GraphDataset_v2, 'sbm': REF.data.GraphDataset_sbm, 'ogb': REF.data.GraphDataset_ogb}[cfg['kind']]
    ds = cls(datas, n_tags=cfg['n_tags'], degree=True)                  # compute_degree + one_hot: reference code
    if graphs[0].get('pe') is not None:
        ds.pe_list = [torch.from_numpy(g['pe']) for g in graphs]
    if graphs[0].get('lap_pe') is not None:
        ds.lap_pe_list = [torch.from_numpy(g['lap_pe']) for g in graphs]
    return ds


def _compact_graphs(graphs, datas=None):
    out = []
    for i, g in enumerate(graphs):
        n = g['x'].shape[0]
        d = dict(x=torch.from_numpy(np.asarray(g['x'])), y=torch.as_tensor(g['y']),
                 edge_index=torch.from_numpy(np.asarray(g['edge_index']).astype(np.int16 if n < 32000 else np.int32)),
                 pe=None if g.get('pe') is None else torch.from_numpy(g['pe']),
                 lap_pe=None if g.get('lap_pe') is None else torch.from_numpy(g['lap_pe']))
        if datas is not None and datas[i].edge_attr is not None:
            d['edge_attr'] = datas[i].edge_attr
        out.append(d)
    return out


def collate_cases():
    out = {}
    for name, n, ids in [("MUTAG", 7, [3, 0, 5, 6]), ("ZINC", 6, [5, 1, 2, 4, 0]), ("PATTERN", 3, [2, 0]),
                         ("CLUSTER", 3, [1, 2, 0]), ("MOLHIV", 6, [0, 4, 2, 5, 1])]:
        graphs = synthetic.make_dataset(name, n, seed=40)
        ds = _ref_dataset(name, graphs)
        batch = ds.collate_fn()([ds[i] for i in ids])
        out[name] = dict(graphs=_compact_graphs(graphs, ds.dataset), ids=ids, batch=list(batch))
    return out


# ---------------------------------------------------------------------------------------------------
# whole models at the five BASELINE shapes (full d=64, reference hyper-parameters, literal all-pairs GCN)
# ---------------------------------------------------------------------------------------------------
def _loss(name, out, labels):
    if name in ("PATTERN", "CLUSTER", "MUTAG"):
        return F.cross_entropy(out, labels.long())
    if name == "ZINC":
        return F.l1_loss(out, labels.to(out.dtype))
    return F.binary_cross_entropy_with_logits(out.reshape(-1), labels.reshape(-1).to(out.dtype))


def _pick(name, want, seed):
    """``want`` graphs of the shape, chosen so SBM batches hold one small and one >128-node graph."""
    pool = synthetic.make_dataset(name, 24, seed=seed)
    if name in ("PATTERN", "CLUSTER"):
        pool.sort(key=lambda g: g['x'].shape[0])
        return [pool[0], pool[-1]][:want] if want <= 2 else [pool[0], pool[len(pool) // 2], pool[-1]]
    return pool[:want]


def model_case(name, B, seed, tag=None, **over):
    cfg = dict(synthetic.CONFIGS[name])
    cfg.update(over)
    graphs = _pick(name, B, seed)
    ds = _ref_dataset(name, graphs)
    batch = ds.collate_fn()([ds[i] for i in range(len(graphs))])                     # reference collate
    px, mask, pe, lap, deg, labels, ei, bi, fi = batch[:9]
    d = cfg['d_model']
    kw = dict(in_size=cfg['n_tags'] if cfg['n_tags'] else 9, nb_class=cfg['nb_class'], d_model=d,
              nb_heads=cfg['heads'], dim_feedforward=2 * d, dropout=0.0, nb_layers=cfg['layers'],
              batch_norm=cfg['batch_norm'], lap_pos_enc=cfg['lap_dim'] > 0, lap_pos_enc_dim=cfg['lap_dim'])
    for k in ('gnn_type', 'last_layer_filter', 'learn_only_filter_order_coeff'):
        if k in cfg:
            kw[k] = cfg[k]
    cls = {'graph': REF.models.DiffGraphTransformerGenGCN, 'node': REF.models.DiffGraphTransformerGenGCNSBM,
           'molhiv': REF.models.DiffGraphTransformerGenGCNMolHiv}[cfg['head']]
    with f64():
        m = cls(**kw)
        det_init(m, seed)
        m.train()
        dbl = lambda t: None if t is None else t.double()
        res = m(dbl(px), ei, bi, fi, mask, dbl(pe), dbl(lap), dbl(deg), return_filter_coeff=True)
        out, coeff = res[0], res[-1]
        loss = _loss(name, out, labels)
        loss.backward()
    grads = {}
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        grads[k] = grad_summary(p.grad) if (p.grad.numel() > BIG_GRAD and p.grad.dim() == 2) else p.grad.clone()
    fix = dict(name=name, seed=seed, over=over, kw=kw, graphs=_compact_graphs(graphs), batch=list(batch[:9]),
               out=out.detach(), coeff=coeff.detach(), loss=loss.detach(), grads=grads,
               no_grad=[k for k, p in m.named_parameters() if p.grad is None])
    save_fixture(fix, os.path.join(HERE, "ref_model_%s.pt.gz" % (tag or name)))
    print("  %-22s B=%d nodes=%s loss=%.6f out|max|=%.4f" % (tag or name, len(graphs),
          [g['x'].shape[0] for g in graphs], float(loss), float(out.abs().max())))


# ---------------------------------------------------------------------------------------------------
# N3: position encodings -- transformer/position_encoding.py:55-161 executed by the reference (scipy expm,
# sparse matrix powers, np.linalg.eig); fp64 where the reference computes in fp64 (it builds fp32 Laplacians:
# get_laplacian's default dtype), stored as the reference returns them
# ---------------------------------------------------------------------------------------------------
def pe_graphs():
    """5 molecule-shape + 2 small dense SBM graphs (the fixture stores dense [n, n] kernels: kept small); #1 gets an
    isolated node, #2 a one-way (directed) edge; every graph also carries symmetric positive edge weights for the
    ``use_edge_attr`` cases."""
    rng = np.random.default_rng(61)
    gs = [synthetic.make_graph(rng, 'ZINC') for _ in range(5)]
    gs = [dict(x=g['x'], edge_index=np.asarray(g['edge_index'])) for g in gs]
    for sizes, p, q in (([7, 9, 8], 0.5, 0.35), ([6, 5, 9, 7], 0.55, 0.25)):
        ei, block = synthetic.sbm_graph(rng, sizes, p, q)
        gs.append(dict(x=block.reshape(-1, 1), edge_index=np.asarray(ei)))
    ei = gs[1]['edge_index']
    gs[1]['edge_index'] = ei[:, (ei[0] != 0) & (ei[1] != 0)]
    n2 = gs[2]['x'].shape[0]
    gs[2]['edge_index'] = np.concatenate([gs[2]['edge_index'], np.array([[0], [n2 - 1]])], axis=1)
    for g in gs:
        s, t = g['edge_index']
        key = np.minimum(s, t) * 100000 + np.maximum(s, t)
        w = {k: rng.uniform(0.5, 2.0) for k in np.unique(key)}
        g['edge_attr'] = np.array([w[k] for k in key], dtype=np.float32)
    return gs


def pe_cases():
    P = REF.pe
    gs = pe_graphs()
    datas = [REF.Data(torch.from_numpy(np.asarray(g['x'])), torch.from_numpy(g['edge_index']).long(), None,
                      torch.from_numpy(g['edge_attr'])) for g in gs]
    encs = {}
    for norm in (None, 'sym', 'rw'):
        tag = str(norm)
        encs['diffusion_' + tag] = P.DiffusionEncoding(None, beta=1.0, normalization=norm)
        encs['diffusion_w_' + tag] = P.DiffusionEncoding(None, beta=0.5, use_edge_attr=True, normalization=norm)
        encs['pstep_' + tag] = P.PStepRWEncoding(None, p=3, beta=0.5, normalization=norm)
        encs['pstep_w_' + tag] = P.PStepRWEncoding(None, p=2, beta=0.25, use_edge_attr=True, normalization=norm)
        encs['lap_' + tag] = P.LapEncoding(4, normalization=norm)
        encs['lap_w_' + tag] = P.LapEncoding(3, use_edge_attr=True, normalization=norm)
    encs['adj'] = P.AdjEncoding(None)
    encs['full'] = P.FullEncoding(None)
    out = dict(graphs=[dict(x=torch.from_numpy(np.asarray(g['x'])), edge_index=torch.from_numpy(g['edge_index']),
                            edge_attr=torch.from_numpy(g['edge_attr'])) for g in gs], pe={})
    for k, enc in encs.items():
        out['pe'][k] = [enc.compute_pe(d) for d in datas]
    return out


if __name__ == "__main__":
    assert ref_shim.available(), "needs the reference tree at %s" % ref_shim.REFERENCE_ROOT
    if sys.argv[1:] == ["pe"]:                       # only the position-encoding fixture
        save_fixture(pe_cases(), os.path.join(HERE, "ref_pe.pt.gz"))
        print("ref_pe.pt.gz %d bytes" % os.path.getsize(os.path.join(HERE, "ref_pe.pt.gz")))
        sys.exit(0)
    save_fixture(pe_cases(), os.path.join(HERE, "ref_pe.pt.gz"))
    coeff, gavg = coeff_cases()
    save_fixture(dict(cheb=cheb_cases(), arma=arma_cases(), coeff=coeff, global_avg=gavg),
                 os.path.join(HERE, "ref_ops.pt.gz"))
    save_fixture(collate_cases(), os.path.join(HERE, "ref_collate.pt.gz"))
    model_case("MUTAG", 5, 51)
    model_case("ZINC", 6, 52)
    model_case("PATTERN", 2, 53)
    model_case("CLUSTER", 2, 54)
    model_case("MOLHIV", 6, 55)
    model_case("ZINC", 5, 56, tag="ZINC_bn", batch_norm=True, layers=3)
    model_case("MUTAG", 4, 57, tag="MUTAG_all_layers", last_layer_filter=False)
    model_case("MUTAG", 4, 58, tag="MUTAG_learn_only", learn_only_filter_order_coeff=True)
    model_case("ZINC", 4, 59, tag="ZINC_arma", gnn_type='ARMAConvDynamic', layers=3)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("ref_") and f.endswith(".pt.gz"):
            print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))
