"""Synthetic graph generators of the BASELINE shapes.

There is no network and the reference ships no datasets, so every benchmark and
parity run uses graphs drawn here.  Shape statistics (nodes / edges / feature vocabularies) are the
public dataset figures quoted in SURVEY.md section 8; weights are random-init.  Each graph is a
dict ``{x, edge_index, y, pe, lap_pe, degree}`` consumable by ``data.GraphStore``.
"""
import numpy as np

from .data import degree_scaling


def _sym_lap(edge_index, n):
    """Dense L_sym = I - D^-1/2 A D^-1/2 (PyG get_laplacian('sym') densified)."""
    A = np.zeros((n, n), dtype=np.float64)
    s, t = edge_index
    keep = s != t
    np.add.at(A, (s[keep], t[keep]), 1.0)
    deg = A.sum(axis=1)
    dis = np.where(deg > 0, 1.0 / np.sqrt(np.maximum(deg, 1e-300)), 0.0)
    return np.eye(n) - dis[:, None] * A * dis[None, :]


def diffusion_pe(edge_index, n, beta=1.0):
    """expm(-beta L_sym) by eigendecomposition (position_encoding.py:65-72)."""
    w, U = np.linalg.eigh(_sym_lap(edge_index, n))
    return ((U * np.exp(-beta * w)[None, :]) @ U.T).astype(np.float32)


def lap_pe(edge_index, n, dim):
    """first `dim` non-trivial eigenvectors of L_sym, zero padded (position_encoding.py:127-161)."""
    w, U = np.linalg.eigh(_sym_lap(edge_index, n))
    pe = U[:, 1:dim + 1]
    out = np.zeros((n, dim), dtype=np.float32)
    out[:, :pe.shape[1]] = pe
    return out


def _undirected(pairs):
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    return np.concatenate([pairs.T, pairs[:, ::-1].T], axis=1)


def molecule_graph(rng, n, extra):
    """random tree + `extra` ring closures -> ~ (n - 1 + extra) undirected bonds."""
    pairs = [(i, int(rng.integers(max(0, i - 3), i))) for i in range(1, n)]
    have = set(map(tuple, map(sorted, pairs)))
    tries = 0
    while extra > 0 and tries < 20 * (extra + 1) and n > 2:
        a, b = sorted(map(int, rng.integers(0, n, size=2)))
        tries += 1
        if a != b and (a, b) not in have:
            have.add((a, b))
            pairs.append((a, b))
            extra -= 1
    return _undirected(pairs)


def sbm_graph(rng, sizes, p, q):
    n = int(np.sum(sizes))
    block = np.repeat(np.arange(len(sizes)), sizes)
    same = block[:, None] == block[None, :]
    prob = np.where(same, p, q)
    upper = np.triu(rng.random((n, n)) < prob, k=1)
    a, b = np.nonzero(upper)
    return _undirected(np.stack([a, b], axis=1)), block


def make_graph(rng, shape, pos_enc=None, lap_dim=0, beta=1.0):
    if shape == 'MUTAG':
        n = int(np.clip(round(rng.normal(17.9, 4.6)), 10, 28))
        ei = molecule_graph(rng, n, max(0, int(round(0.1 * n))) + 1)
        g = dict(x=rng.integers(0, 7, size=(n, 1)), edge_index=ei, y=np.int64(rng.integers(0, 2)))
    elif shape == 'ZINC':
        n = int(np.clip(round(rng.normal(23.2, 4.5)), 9, 37))
        ei = molecule_graph(rng, n, int(rng.integers(2, 4)))
        g = dict(x=rng.integers(0, 28, size=(n, 1)), edge_index=ei,
                 y=np.array([rng.normal()], dtype=np.float32))
    elif shape == 'MOLHIV':
        n = int(np.clip(round(rng.lognormal(np.log(25.5) - 0.5 * 0.45 ** 2, 0.45)), 2, 222))
        ei = molecule_graph(rng, n, max(0, int(round(0.08 * n))) + 1)
        vocab = [119, 4, 12, 12, 10, 6, 6, 2, 2]
        x = np.stack([rng.integers(0, v, size=n) for v in vocab], axis=1)
        g = dict(x=x.astype(np.float32), edge_index=ei, y=np.array([rng.integers(0, 2)], dtype=np.float32))
    elif shape in ('PATTERN', 'CLUSTER'):
        if shape == 'PATTERN':
            sizes = list(rng.integers(5, 36, size=5)) + [20]
            ei, block = sbm_graph(rng, sizes, 0.5, 0.35)
            y = (block == 5).astype(np.int64)
            x = rng.integers(0, 3, size=(len(block), 1))
        else:
            sizes = list(rng.integers(5, 36, size=6))
            ei, block = sbm_graph(rng, sizes, 0.55, 0.25)
            y = block.astype(np.int64)
            x = rng.integers(0, 7, size=(len(block), 1))
        g = dict(x=x, edge_index=ei, y=y)
    else:
        raise ValueError(shape)
    n = g['x'].shape[0]
    g['degree'] = degree_scaling(g['edge_index'], n)
    g['pe'] = diffusion_pe(g['edge_index'], n, beta) if pos_enc == 'diffusion' else None
    g['lap_pe'] = lap_pe(g['edge_index'], n, lap_dim) if lap_dim > 0 else None
    return g


# the five BASELINE configs: model hyper-parameters follow the reference drivers' defaults
CONFIGS = {
    # experiments/run_transformer_gengcn_cv.py:34-36,51
    'MUTAG': dict(shape='MUTAG', kind='v2', n_tags=7, nb_class=2, batch=32, heads=4, layers=3, d_model=64,
                  pos_enc=None, lap_dim=0, batch_norm=False, head='graph'),
    # experiments/run_transformer_gengcn.py:35-37,52 (+ --pos-enc diffusion --beta 1.0; LN variant)
    'ZINC': dict(shape='ZINC', kind='v2', n_tags=28, nb_class=1, batch=128, heads=8, layers=10, d_model=64,
                 pos_enc='diffusion', lap_dim=0, batch_norm=False, head='graph'),
    # experiments/run_transformer_gengcn_SBM_cv.py:36-38 with --batch-size 64
    'PATTERN': dict(shape='PATTERN', kind='sbm', n_tags=3, nb_class=2, batch=64, heads=4, layers=3, d_model=64,
                    pos_enc=None, lap_dim=0, batch_norm=False, head='node'),
    'CLUSTER': dict(shape='CLUSTER', kind='sbm', n_tags=7, nb_class=6, batch=64, heads=4, layers=3, d_model=64,
                    pos_enc=None, lap_dim=8, batch_norm=False, head='node'),
    # experiments/run_transformer_gengcn_molhiv.py:43-45 with --batch-size 1024
    'MOLHIV': dict(shape='MOLHIV', kind='ogb', n_tags=None, nb_class=1, batch=1024, heads=4, layers=3,
                   d_model=64, pos_enc=None, lap_dim=0, batch_norm=False, head='molhiv'),
}


def make_dataset(name, num_graphs, seed=0):
    cfg = CONFIGS[name]
    rng = np.random.default_rng(seed)
    return [make_graph(rng, cfg['shape'], pos_enc=cfg['pos_enc'], lap_dim=cfg['lap_dim'])
            for _ in range(num_graphs)]


def build_model(name, module, **overrides):
    """Instantiate the config's model from ``module`` (``feta_tmlr_b200.models`` or, in tests and
    the CPU baseline, the oracle's model module -- passed in by the caller)."""
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    d = cfg['d_model']
    kw = dict(in_size=cfg['n_tags'] if cfg['n_tags'] else 9, nb_class=cfg['nb_class'], d_model=d,
              nb_heads=cfg['heads'], dim_feedforward=2 * d, dropout=0.0, nb_layers=cfg['layers'],
              batch_norm=cfg['batch_norm'], lap_pos_enc=cfg['lap_dim'] > 0, lap_pos_enc_dim=cfg['lap_dim'])
    names = {'graph': ('DiffGraphTransformerGenGCN', 'OracleDiffGraphTransformerGenGCN'),
             'node': ('DiffGraphTransformerGenGCNSBM', 'OracleDiffGraphTransformerGenGCNSBM'),
             'molhiv': ('DiffGraphTransformerGenGCNMolHiv', 'OracleDiffGraphTransformerGenGCNMolHiv')}[cfg['head']]
    cls = getattr(module, names[0], None) or getattr(module, names[1])
    for k in ('gnn_type', 'last_layer_filter', 'learn_only_filter_order_coeff'):
        if k in cfg:
            kw[k] = cfg[k]
    return cls(**kw)
