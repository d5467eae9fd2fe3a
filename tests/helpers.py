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
