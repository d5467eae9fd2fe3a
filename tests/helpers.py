"""Shared builders for the parity tests (seeded inputs, oracle <-> product plumbing)."""
import numpy as np
import torch

import oracle.data as od
from feta_tmlr_b200 import data as fdata
from feta_tmlr_b200 import synthetic


def random_batch_graph(seed, sizes, deg=3.0, self_loops=True, multi_edges=True, directed_extra=0):
    """Block-diagonal random edge list over graphs of the given sizes (sizes may include 1)."""
    rng = np.random.default_rng(seed)
    src, dst, off = [], [], 0
    for n in sizes:
        m = int(round(deg * n / 2))
        if n > 1 and m > 0:
            a = rng.integers(0, n, size=m)
            b = rng.integers(0, n, size=m)
            keep = a != b
            a, b = a[keep], b[keep]
            src += list(a + off) + list(b + off)
            dst += list(b + off) + list(a + off)
            if multi_edges and len(a):
                src += [a[0] + off, b[0] + off]
                dst += [b[0] + off, a[0] + off]
            for _ in range(directed_extra):
                u, v = rng.integers(0, n, size=2)
                src.append(u + off)
                dst.append(v + off)
        if self_loops and n > 0:
            src.append(off)
            dst.append(off)
        off += n
    ei = torch.tensor(np.array([src, dst], dtype=np.int64).reshape(2, -1))
    batch = torch.tensor(np.repeat(np.arange(len(sizes)), sizes), dtype=torch.int64)
    return ei, batch, off


def oracle_csr(edge_index, R, lambda_max=2.0, transpose=False):
    """CSR the product must match: L_hat entries grouped by target (or source), input order."""
    s, t = edge_index[0].numpy(), edge_index[1].numpy()
    keep = s != t
    s, t = s[keep], t[keep]
    deg = np.bincount(s, minlength=R).astype(np.float32)
    with np.errstate(divide='ignore'):
        dis = np.where(deg > 0, (1.0 / np.sqrt(deg)).astype(np.float32), np.float32(0))
    w = (-(dis[s] * dis[t]) * np.float32(2.0 / lambda_max)).astype(np.float32)
    key, other = (s, t) if transpose else (t, s)
    order = np.argsort(key, kind='stable')
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=R))]).astype(np.int32)
    return rowptr, other[order].astype(np.int32), w[order]


def oracle_graphs(graphs, n_tags):
    out = []
    for g in graphs:
        og = od.Graph(torch.from_numpy(np.asarray(g['x'])), torch.from_numpy(g['edge_index']),
                      torch.as_tensor(g['y']))
        if n_tags:
            og.x_onehot = od.one_hot(og, n_tags)
        og.degree = od.compute_degree(og)
        if g.get('pe') is not None:
            og.pe = torch.from_numpy(g['pe'])
        if g.get('lap_pe') is not None:
            og.lap_pe = torch.from_numpy(g['lap_pe'])
        out.append(og)
    return out


def make_batch(name, num_graphs, seed, ids=None):
    cfg = synthetic.CONFIGS[name]
    graphs = synthetic.make_dataset(name, num_graphs, seed=seed)
    store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
    ids = np.arange(num_graphs) if ids is None else np.asarray(ids)
    return cfg, graphs, store, fdata.collate_host(store, ids)


def to_dev(batch, dev):
    return tuple(None if t is None else t.to(dev) for t in batch)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------------
# reference-generated fixtures (tests/golden/make_golden_from_reference.py)
# ---------------------------------------------------------------------------------------------------
def det_init(module, seed):
    """Deterministic, library-independent parameter values (NumPy legacy MT19937 stream, which NumPy keeps
    frozen): glorot-uniform over the last two dims for matrices, 1 +- 0.1 for norm scales, +-0.1 for every
    other vector.  The generator and the tests both call this, so fixtures never store weights."""
    import math
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            shape = tuple(p.shape)
            if p.dim() >= 2:
                a = math.sqrt(6.0 / (shape[-2] + shape[-1]))
                v = rs.uniform(-a, a, size=shape)
            elif 'norm' in name and name.endswith('weight'):
                v = 1.0 + 0.1 * rs.uniform(-1, 1, size=shape)
            else:
                v = 0.1 * rs.uniform(-1, 1, size=shape)
            p.copy_(torch.from_numpy(v.astype(np.float32)).to(p.dtype))
    return module


BIG_GRAD = 1 << 12          # gradients with more entries than this are stored as summaries


def grad_summary(g):
    """Compact stand-in for a large 2-D gradient: row sums, column sums, 4 rows and 4 columns."""
    g = g.detach()
    r = torch.linspace(0, g.shape[0] - 1, 4).long()
    c = torch.linspace(0, g.shape[1] - 1, 4).long()
    return dict(summary=True, shape=tuple(g.shape), rowsum=g.sum(1), colsum=g.sum(0), rows=g[r], cols=g[:, c],
                absmax=g.abs().max())


def grad_scale(wants):
    """Largest gradient entry over a fixture's whole gradient dict (tensors or summaries)."""
    return max([float(w['absmax'] if isinstance(w, dict) else w.abs().max()) for w in wants.values()] + [0.0])


def check_grad(got, want, tol, name="", floor=0.0):
    """``want``: a tensor or a ``grad_summary`` dict (fixture side).  Error relative to the largest entry of
    THIS gradient, or to ``floor`` when the whole gradient is smaller than that (a gradient that is analytically
    zero -- e.g. a bias followed by BatchNorm -- is rounding noise on both sides)."""
    got = got.detach().double().cpu()
    if isinstance(want, dict):
        s = grad_summary(got)
        for k in ("rowsum", "colsum", "rows", "cols"):
            ref = want[k].double()
            scale = max(float(ref.abs().max()), float(want['absmax']), floor, 1e-30)     # a summary can be all ~0
            err = float((s[k].double() - ref).abs().max()) / scale
            assert err < tol, (name, k, err)
        return
    want = want.double()
    err = float((got - want).abs().max()) / max(float(want.abs().max()), floor, 1e-30)
    assert err < tol, (name, err)


def graphs_from_fixture(glist):
    """Stored compact graphs -> the dict form synthetic.make_graph produces."""
    out = []
    for g in glist:
        d = dict(x=g['x'].numpy(), edge_index=g['edge_index'].numpy().astype(np.int64), y=g['y'].numpy())
        n = d['x'].shape[0]
        d['degree'] = fdata.degree_scaling(d['edge_index'], n)
        d['pe'] = g['pe'].numpy() if g.get('pe') is not None else None
        d['lap_pe'] = g['lap_pe'].numpy() if g.get('lap_pe') is not None else None
        if g.get('edge_attr') is not None:
            d['edge_attr'] = g['edge_attr'].numpy()
        out.append(d)
    return out


def pack_tree(obj, f32=False):
    """Shrink a fixture tree for storage: integer tensors go to the narrowest dtype that holds them (original
    dtype recorded), fp64 tensors go to fp32 when ``f32`` (used for gradients: 6e-8 rounding vs a 1e-4 gate)."""
    if torch.is_tensor(obj):
        if obj.dtype in (torch.int64, torch.int32):
            lo, hi = (int(obj.min()), int(obj.max())) if obj.numel() else (0, 0)
            for dt, bound in ((torch.int8, 127), (torch.int16, 32767), (torch.int32, 2 ** 31 - 1)):
                if -bound <= lo and hi <= bound and dt != obj.dtype:
                    return dict(__packed__=str(obj.dtype), data=obj.to(dt))
            return obj
        if f32 and obj.dtype == torch.float64:
            return obj.float()
        return obj.clone()
    if isinstance(obj, dict):
        return {k: pack_tree(v, f32 or k in ('grads', 'dx', 'dcoeff', 'dbias', 'dweight', 'coeff')) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [pack_tree(v, f32) for v in obj]
    return obj


def unpack_tree(obj):
    if isinstance(obj, dict):
        if '__packed__' in obj:
            return obj['data'].to(getattr(torch, obj['__packed__'].split('.')[-1]))
        return {k: unpack_tree(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [unpack_tree(v) for v in obj]
    return obj


def save_fixture(obj, path):
    import gzip
    import io
    buf = io.BytesIO()
    torch.save(pack_tree(obj), buf)
    with gzip.open(path, 'wb', compresslevel=9) as f:
        f.write(buf.getvalue())


def load_fixture(name):
    import gzip
    import io
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name)
    with gzip.open(path, 'rb') as f:
        return unpack_tree(torch.load(io.BytesIO(f.read()), weights_only=False))
