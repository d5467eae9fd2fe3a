"""Mini-batch construction (reference: transformer/data.py:113-460).

``GraphStore`` packs a dataset once (node features, graph-local edge lists, dense per-graph PE
blocks, degree vectors, labels) with prefix-sum pointers.  Two builders produce the reference's
collate tuple ``(padded_x, mask, pos_enc, lap_pos_enc, degree, labels, edge_indices,
batch_indices, feature_indices_to_gather)`` bit-identically:

  * ``collate_host``   -- vectorised NumPy on the host (no per-node Python loops, data.py:218),
                          the three index tensors moved to the device like data.py:224;
  * ``DeviceBatchBuilder`` -- the store lives in HBM and the batch is assembled by the
                          ``feta_collate_*`` kernels (csrc/collate.cu) from the graph ids alone.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check


def degree_scaling(edge_index, num_nodes):
    """data.py:142-146: 1 / sqrt(1 + degree(edge_index[0]))."""
    deg = np.bincount(np.asarray(edge_index[0]), minlength=num_nodes).astype(np.float32)
    return (1.0 / np.sqrt(1.0 + deg)).astype(np.float32)


class GraphStore(object):
    """Packed dataset.  ``kind`` selects the label convention of the reference's three collates:
    'v2' / 'ogb' stack one label row per graph (default_collate), 'sbm' concatenates per-node
    labels (data.py:457)."""

    def __init__(self, graphs, kind='v2', n_tags=None):
        assert kind in ('v2', 'sbm', 'ogb')
        self.kind, self.n_tags = kind, n_tags
        n_nodes = np.array([int(g['x'].shape[0]) for g in graphs], dtype=np.int64)
        n_edges = np.array([int(g['edge_index'].shape[1]) for g in graphs], dtype=np.int64)
        self.num_graphs = len(graphs)
        self.node_ptr = np.concatenate([[0], np.cumsum(n_nodes)]).astype(np.int64)
        self.edge_ptr = np.concatenate([[0], np.cumsum(n_edges)]).astype(np.int64)
        self.pe_ptr = np.concatenate([[0], np.cumsum(n_nodes * n_nodes)]).astype(np.int64)
        x = np.concatenate([np.asarray(g['x']).reshape(g['x'].shape[0], -1) for g in graphs], axis=0)
        if n_tags is not None and n_tags > 1:                      # data.py:153-159 one_hot
            tags = x.reshape(-1).astype(np.int64)
            onehot = np.zeros((tags.shape[0], n_tags), dtype=np.float32)
            onehot[np.arange(tags.shape[0]), tags] = 1.0
            self.x = onehot
        else:
            self.x = x.astype(np.float32)
        self.n_features = self.x.shape[1]
        self.edge_index = np.concatenate([np.asarray(g['edge_index'], dtype=np.int64) for g in graphs], axis=1)
        self.has_pe = graphs[0].get('pe') is not None
        self.pe = np.concatenate([np.asarray(g['pe'], dtype=np.float32).reshape(-1) for g in graphs]) \
            if self.has_pe else None
        self.has_lap = graphs[0].get('lap_pe') is not None
        self.lap_pe = np.concatenate([np.asarray(g['lap_pe'], dtype=np.float32) for g in graphs], axis=0) \
            if self.has_lap else None
        self.has_degree = graphs[0].get('degree') is not None
        self.degree = np.concatenate([np.asarray(g['degree'], dtype=np.float32) for g in graphs]) \
            if self.has_degree else None
        if kind == 'sbm':
            self.y = np.concatenate([np.asarray(g['y']).reshape(-1) for g in graphs])
        else:
            self.y = np.stack([np.asarray(g['y']) for g in graphs], axis=0)
        self.has_edge_attr = kind == 'ogb' and graphs[0].get('edge_attr') is not None
        self.edge_attr = np.concatenate([np.asarray(g['edge_attr']) for g in graphs], axis=0) \
            if self.has_edge_attr else None

    def sizes(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        return self.node_ptr[ids + 1] - self.node_ptr[ids], self.edge_ptr[ids + 1] - self.edge_ptr[ids]


def _ranges(starts, lens):
    """concatenate [arange(s, s+l) for s, l in zip(starts, lens)] without a Python loop."""
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    out_ptr = np.concatenate([[0], np.cumsum(lens)])[:-1]
    owner = np.repeat(np.arange(len(lens)), lens)
    return np.arange(total, dtype=np.int64) - out_ptr[owner] + starts[owner], owner


def collate_host(store, ids, device='cpu', static=None, edge_dtype=None):
    """Vectorised restatement of GraphDataset_*.collate_fn (data.py:161-225 / :277-344 / :394-460).

    ``static=(nmax_cap, e_cap)`` (extension, for engine.GraphedTrainStep): pad the node axis to
    ``nmax_cap`` instead of the batch maximum, pad ``edge_indices`` to ``e_cap`` columns with
    (-1, -1) columns (ignored by the plan builder), and (node-level labels) return labels as
    ``[B, nmax_cap]`` with -100 in padded slots -- every returned shape but batch_indices /
    feature_indices is then batch independent.  ``edge_dtype=np.int32`` (static only): ship the edge list
    as int32 -- it is ~95 % of a PATTERN-shape mini-batch's host->device bytes -- and let
    ``forward_static`` widen it on the device."""
    ids = np.asarray(ids, dtype=np.int64)
    B = len(ids)
    lens, elens = store.sizes(ids)
    nmax = int(lens.max())
    if static is not None:
        if nmax > static[0] or int(elens.sum()) > static[1]:
            raise ValueError("batch exceeds the static capacity: nmax %d > %d or E %d > %d"
                             % (nmax, static[0], int(elens.sum()), static[1]))
        nmax = int(static[0])
    src_rows, owner = _ranges(store.node_ptr[ids], lens)
    local = np.arange(len(owner), dtype=np.int64) - np.concatenate([[0], np.cumsum(lens)])[:-1][owner]
    padded_x = np.zeros((B, nmax, store.n_features), dtype=np.float32)
    padded_x[owner, local] = store.x[src_rows]
    mask = np.arange(nmax)[None, :] >= lens[:, None]                                 # :172, :210
    pos_enc = None
    if store.has_pe:
        pos_enc = np.zeros((B, nmax, nmax), dtype=np.float32)                        # :182
        for b, gid in enumerate(ids):                                                # per graph block copy
            n = int(lens[b])
            pos_enc[b, :n, :n] = store.pe[store.pe_ptr[gid]:store.pe_ptr[gid + 1]].reshape(n, n)
    lap = None
    if store.has_lap:
        lap = np.zeros((B, nmax, store.lap_pe.shape[1]), dtype=np.float32)           # :191
        lap[owner, local] = store.lap_pe[src_rows]
    degree = None
    if store.has_degree:
        degree = np.zeros((B, nmax), dtype=np.float32)                               # :196
        degree[owner, local] = store.degree[src_rows]
    node_off = np.concatenate([[0], np.cumsum(lens)])[:-1]
    e_src, e_owner = _ranges(store.edge_ptr[ids], elens)
    edge_indices = store.edge_index[:, e_src] + node_off[e_owner][None, :]           # :219
    if static is not None:
        padded_e = np.full((2, int(static[1])), -1, dtype=np.int64)     # (-1, -1): ignored by the plan builder
        padded_e[:, :edge_indices.shape[1]] = edge_indices
        edge_indices = padded_e if edge_dtype is None else padded_e.astype(edge_dtype)
    batch_indices = owner.astype(np.int64)                                           # :220
    feature_indices = np.stack([owner, local], axis=1).astype(np.int64)              # :218
    if store.kind == 'sbm':
        y_rows, _ = _ranges(store.node_ptr[ids], lens)
        if static is not None:
            lab = np.full((B, nmax), -100, dtype=np.int64)
            lab[owner, local] = store.y[y_rows]
            labels = torch.from_numpy(lab)
        else:
            labels = torch.from_numpy(store.y[y_rows])                               # :457
    else:
        labels = torch.from_numpy(store.y[ids])                                      # default_collate
    t = torch.from_numpy
    out = (t(padded_x), t(mask), None if pos_enc is None else t(pos_enc), None if lap is None else t(lap),
           None if degree is None else t(degree), labels, t(edge_indices).to(device),
           t(batch_indices).to(device), t(feature_indices).to(device))              # :224
    if store.kind == 'ogb' and store.has_edge_attr:
        out = out + (t(store.edge_attr[e_src]).to(device),)                          # :342-343
    return out


class DeviceBatchBuilder(object):
    """GPU batch builder: dataset resident in HBM, batches assembled by csrc/collate.cu.

    Only the B graph ids and two (B+1)-entry prefix sums cross PCIe per step."""

    def __init__(self, store, device='cuda'):
        self.store = store
        self.device = torch.device(device)
        d = self.device
        self.node_ptr = torch.from_numpy(store.node_ptr).to(d)
        self.edge_ptr = torch.from_numpy(store.edge_ptr).to(d)
        self.pe_ptr = torch.from_numpy(store.pe_ptr).to(d)
        self.x = torch.from_numpy(store.x).to(d)
        self.edge_index = torch.from_numpy(store.edge_index).contiguous().to(d)
        self.pe = torch.from_numpy(store.pe).to(d) if store.has_pe else None
        self.lap_pe = torch.from_numpy(store.lap_pe).to(d) if store.has_lap else None
        self.degree = torch.from_numpy(store.degree).to(d) if store.has_degree else None
        self.y = torch.from_numpy(store.y).to(d)

    def build(self, ids, static=None):
        """``static=(nmax_cap, e_cap)``: the static-shape tuple of ``collate_host(..., static=...)`` (node axis
        padded to ``nmax_cap``, ``edge_indices [2, e_cap]`` padded with (-1, -1), node-level labels
        ``[B, nmax_cap]`` with -100 in padded slots) so the batch can feed ``engine.GraphedTrainStep``."""
        lib = _lib.load()
        st = self.store
        ids = np.asarray(ids, dtype=np.int64)
        B = len(ids)
        lens, elens = st.sizes(ids)
        nmax, N, E = int(lens.max()), int(lens.sum()), int(elens.sum())
        e_cols = E
        if static is not None:
            if nmax > static[0] or E > static[1]:
                raise ValueError("batch exceeds the static capacity: nmax %d > %d or E %d > %d"
                                 % (nmax, static[0], E, static[1]))
            nmax, e_cols = int(static[0]), int(static[1])
        host = np.concatenate([ids, np.concatenate([[0], np.cumsum(lens)]),
                               np.concatenate([[0], np.cumsum(elens)])]).astype(np.int64)
        dev = torch.from_numpy(host).pin_memory().to(self.device, non_blocking=True)
        gid, onp, oep = dev[:B], dev[B:2 * B + 1], dev[2 * B + 1:]
        d = self.device
        stream = torch.cuda.current_stream().cuda_stream
        mask = torch.empty((B, nmax), dtype=torch.bool, device=d)
        edge_indices = torch.empty((2, e_cols), dtype=torch.int64, device=d)
        batch_indices = torch.empty((N,), dtype=torch.int64, device=d)
        feature_indices = torch.empty((N, 2), dtype=torch.int64, device=d)
        check(lib.feta_collate_indices(gid.data_ptr(), self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(),
                                       self.edge_index.data_ptr(), self.edge_index.shape[1], onp.data_ptr(),
                                       oep.data_ptr(), mask.data_ptr(),
                                       edge_indices.data_ptr() if static is None else 0,
                                       batch_indices.data_ptr(), feature_indices.data_ptr(), B, nmax, N,
                                       E if static is None else 0, stream), "feta_collate_indices")
        if static is not None:
            check(lib.feta_collate_edges_static(gid.data_ptr(), self.edge_ptr.data_ptr(), self.edge_index.data_ptr(),
                                                self.edge_index.shape[1], onp.data_ptr(), oep.data_ptr(),
                                                edge_indices.data_ptr(), B, E, e_cols, stream),
                  "feta_collate_edges_static")

        def pad_rows(src, C):
            dst = torch.empty((B, nmax, C), dtype=torch.float32, device=d)
            check(lib.feta_collate_pad_rows(gid.data_ptr(), self.node_ptr.data_ptr(), src.data_ptr(),
                                            dst.data_ptr(), B, nmax, C, stream), "feta_collate_pad_rows")
            return dst

        padded_x = pad_rows(self.x, st.n_features)
        pos_enc = None
        if self.pe is not None:
            pos_enc = torch.empty((B, nmax, nmax), dtype=torch.float32, device=d)
            check(lib.feta_collate_pad_pe(gid.data_ptr(), self.node_ptr.data_ptr(), self.pe_ptr.data_ptr(),
                                          self.pe.data_ptr(), pos_enc.data_ptr(), B, nmax, stream),
                  "feta_collate_pad_pe")
        lap = pad_rows(self.lap_pe, self.lap_pe.shape[1]) if self.lap_pe is not None else None
        degree = pad_rows(self.degree.view(-1, 1), 1).view(B, nmax) if self.degree is not None else None
        if st.kind == 'sbm':
            rows, _ = _ranges(st.node_ptr[ids], lens)
            labels = self.y[torch.from_numpy(rows).to(d)]
            if static is not None:
                lab = torch.full((B, nmax), -100, dtype=torch.int64, device=d)
                lab[feature_indices[:, 0], feature_indices[:, 1]] = labels
                labels = lab
        else:
            labels = self.y[gid]
        return (padded_x, mask, pos_enc, lap, degree, labels, edge_indices, batch_indices, feature_indices)
