"""Position encodings (reference: transformer/position_encoding.py:11-176) -- SURVEY.md section 8(f) N3.

Same class names, constructor arguments, ``apply_to`` behaviour and pickle cache format
(``<savepath>.<split>`` holding a list of dense per-graph tensors, :35-49) as the reference, so a
cache written by either side is readable by the other.  What changes is how the kernels are
computed: the reference calls ``scipy.sparse.linalg.expm`` / ``np.linalg.eig`` once per graph on
the host (minutes on full ZINC / molhiv); here graphs are grouped into size-sorted chunks, each padded into
a ``[G, n, n]`` batch (<= 256 MB) and decomposed with one batched symmetric eigendecomposition
(``torch.linalg.eigh``, on the GPU when one is present):

    L_sym = U diag(w) U^T   ->   expm(-beta L) = U diag(exp(-beta w)) U^T        (DiffusionEncoding)
                                 (I - beta L)^p = U diag((1 - beta w)^p) U^T     (PStepRWEncoding)
                                 columns 1..dim of U (ascending w)               (LapEncoding)

``'sym'`` and ``None`` normalisations are symmetric and take the batched path directly.  ``'rw'`` is not
symmetric but, for an undirected edge list, similar to ``'sym'``:  ``L_rw = S^-1 L_sym S`` with
``S = diag(sqrt(deg))`` (1 for isolated nodes, which are decoupled 1x1 blocks), so

    f(L_rw) = S^-1 U f(w) U^T S        and the eigenvectors of L_rw are  S^-1 u / |S^-1 u|

ride on the same batched ``eigh``.  A directed edge list (non-symmetric adjacency) takes the general per-graph
dense formula instead.  Graphs are dicts / objects with ``edge_index`` and ``num_nodes`` (``x.shape[0]``), as
produced by ``synthetic`` or PyG.
"""
import os
import pickle

import numpy as np
import torch


def _num_nodes(g):
    if isinstance(g, dict):
        return int(g.get('num_nodes', g['x'].shape[0]))
    return int(getattr(g, 'num_nodes', None) or g.x.shape[0])


def _edge_index(g):
    ei = g['edge_index'] if isinstance(g, dict) else g.edge_index
    return np.asarray(ei.cpu() if torch.is_tensor(ei) else ei, dtype=np.int64)


def _edge_weight(g):
    """``graph.edge_attr`` as the 1-D edge weight PyG ``get_laplacian`` takes (``use_edge_attr=True``, :66)."""
    ea = g['edge_attr'] if isinstance(g, dict) else g.edge_attr
    return np.asarray(ea.cpu() if torch.is_tensor(ea) else ea, dtype=np.float64).reshape(-1)


def dense_laplacians(graphs, normalization, device='cpu', dtype=torch.float64, return_deg=False, weighted=False):
    """Padded batch ``[G, nmax, nmax]`` of PyG ``get_laplacian`` matrices (self loops removed,
    multi-edges summed, degree over the source index; ``weighted``: edge weights from ``edge_attr``, assumed
    positive) + the node counts (+ the degrees)."""
    ns = np.array([_num_nodes(g) for g in graphs], dtype=np.int64)
    nmax = int(ns.max())
    A = np.zeros((len(graphs), nmax, nmax), dtype=np.float64)
    for i, g in enumerate(graphs):
        s, t = _edge_index(g)
        keep = s != t
        np.add.at(A[i], (s[keep], t[keep]), _edge_weight(g)[keep] if weighted else 1.0)
    A = torch.from_numpy(A).to(device=device, dtype=dtype)
    deg = A.sum(dim=2)
    eye = torch.eye(nmax, device=device, dtype=dtype).unsqueeze(0)
    real = (torch.arange(nmax, device=device).unsqueeze(0) < torch.from_numpy(ns).to(device).unsqueeze(1))
    real2 = (real.unsqueeze(1) & real.unsqueeze(2)).to(dtype)
    if normalization is None:
        L = torch.diag_embed(deg) - A
    elif normalization == 'sym':
        dis = torch.where(deg > 0, deg.clamp(min=1e-300).rsqrt(), torch.zeros_like(deg))
        L = eye * real2 - dis.unsqueeze(2) * A * dis.unsqueeze(1)
    elif normalization == 'rw':
        di = torch.where(deg > 0, 1.0 / deg.clamp(min=1e-300), torch.zeros_like(deg))
        L = eye * real2 - di.unsqueeze(2) * A
    else:
        raise ValueError("normalization must be None, 'sym' or 'rw'")
    if return_deg:
        return L, ns, deg
    return L, ns


CHUNK_BYTES = 256 << 20      # padded fp64 batch per eigh call (L, U and the product are each this big)


def _size_sorted_chunks(graphs):
    """Index chunks of similar-size graphs, each padded batch <= CHUNK_BYTES (the whole dataset in one padded
    batch would be ~16 GB per copy for molhiv: 41k graphs padded to 222 nodes)."""
    ns = np.array([_num_nodes(g) for g in graphs], dtype=np.int64)
    order = np.argsort(ns, kind='stable')
    i = 0
    while i < len(order):
        j = i + 1
        while j < len(order) and (j + 1 - i) * int(ns[order[j]]) ** 2 * 8 <= CHUNK_BYTES:
            j += 1
        yield order[i:j]
        i = j


def _is_symmetric(L):
    return (L == L.transpose(1, 2)).flatten(1).all(dim=1)


def _rw_scale(deg):
    """S of ``L_rw = S^-1 L_sym S``: sqrt(deg), 1 where the degree is 0 (isolated and padded nodes)."""
    return torch.where(deg > 0, deg.clamp(min=1e-300).sqrt(), torch.ones_like(deg))


def _spectral_map(graphs, normalization, fn, device, fallback, weighted=False):
    """U f(w) U^T per graph, one batched eigh per size-sorted chunk (``'rw'``: the similarity transform of the
    module docstring around the ``'sym'`` decomposition).  A graph whose adjacency is NOT symmetric (directed edge
    list) takes ``fallback(L)`` -- the general dense formula the reference's scipy ``expm`` / matrix power
    computes -- instead of a silently wrong ``eigh``."""
    out = [None] * len(graphs)
    rw = normalization == 'rw'
    for idx in _size_sorted_chunks(graphs):
        sub = [graphs[i] for i in idx]
        L, ns, deg = dense_laplacians(sub, 'sym' if rw else normalization, device=device, return_deg=True,
                                      weighted=weighted)
        sym = _is_symmetric(L).cpu().numpy()
        # padded rows/cols are zero: they add zero eigenvalues whose eigenvectors live in the padding
        w, U = torch.linalg.eigh(L)
        M = (U * fn(w).unsqueeze(1)) @ U.transpose(1, 2)
        if rw:
            s = _rw_scale(deg)
            M = M * s.unsqueeze(1) / s.unsqueeze(2)
        for j, (i, n) in enumerate(zip(idx, ns)):
            if sym[j]:
                out[i] = M[j, :n, :n].to(torch.float32).cpu()
            else:
                Lj = dense_laplacians([sub[j]], normalization, device=device, weighted=weighted)[0][0] if rw \
                    else L[j, :n, :n]
                out[i] = fallback(Lj).to(torch.float32).cpu()
    return out


class PositionEncoding(object):
    """transformer/position_encoding.py:11-52."""

    def __init__(self, savepath=None, zero_diag=False, device=None):
        self.savepath = savepath
        self.zero_diag = zero_diag
        self.device = device or ('cuda' if torch.cuda.is_available() else 'cpu')

    def apply_to(self, dataset, split='train'):
        saved_pos_enc = self.load(split)
        graphs = [dataset[i] for i in range(len(dataset))]
        all_pe = saved_pos_enc if saved_pos_enc is not None else self.compute_all(graphs)
        dataset.pe_list = []
        for pe in all_pe:
            if self.zero_diag:
                pe = pe.clone()
                pe.diagonal()[:] = 0
            dataset.pe_list.append(pe)
        if saved_pos_enc is None:
            self.save(all_pe, split)
        return dataset

    def save(self, pos_enc, split):
        if self.savepath is None:
            return
        if not os.path.isfile(self.savepath + "." + split):
            with open(self.savepath + "." + split, 'wb') as handle:
                pickle.dump(pos_enc, handle)

    def load(self, split):
        if self.savepath is None:
            return None
        if not os.path.isfile(self.savepath + "." + split):
            return None
        with open(self.savepath + "." + split, 'rb') as handle:
            return pickle.load(handle)

    def compute_all(self, graphs):
        return [self.compute_pe(g) for g in graphs]

    def compute_pe(self, graph):
        return self.compute_all([graph])[0]


class DiffusionEncoding(PositionEncoding):
    """:55-72 -- expm(-beta L)."""

    def __init__(self, savepath, beta=1., use_edge_attr=False, normalization=None, zero_diag=False, device=None):
        super().__init__(savepath, zero_diag, device)
        self.beta, self.normalization, self.use_edge_attr = beta, normalization, use_edge_attr

    def compute_all(self, graphs):
        return _spectral_map(graphs, self.normalization, lambda w: torch.exp(-self.beta * w), self.device,
                             fallback=lambda L: torch.matrix_exp(-self.beta * L), weighted=self.use_edge_attr)


class PStepRWEncoding(PositionEncoding):
    """:75-93 -- (I - beta L)^p."""

    def __init__(self, savepath, p=1, beta=0.5, use_edge_attr=False, normalization=None, zero_diag=False,
                 device=None):
        super().__init__(savepath, zero_diag, device)
        self.p, self.beta, self.normalization, self.use_edge_attr = p, beta, normalization, use_edge_attr

    def compute_all(self, graphs):
        return _spectral_map(graphs, self.normalization, lambda w: (1.0 - self.beta * w) ** self.p, self.device,
                             fallback=lambda L: torch.linalg.matrix_power(
                                 torch.eye(L.shape[0], dtype=L.dtype, device=L.device) - self.beta * L, self.p),
                             weighted=self.use_edge_attr)


class AdjEncoding(PositionEncoding):
    """:96-105 -- dense adjacency ([1, n, n] like PyG ``to_dense_adj``)."""

    def __init__(self, savepath, normalization=None, zero_diag=False, device=None):
        super().__init__(savepath, zero_diag, device)
        self.normalization = normalization

    def compute_all(self, graphs):
        out = []
        for g in graphs:
            n = _num_nodes(g)
            A = np.zeros((n, n), dtype=np.float32)
            s, t = _edge_index(g)
            np.add.at(A, (s, t), 1.0)
            out.append(torch.from_numpy(A).unsqueeze(0))
        return out


class FullEncoding(PositionEncoding):
    """:107-116 -- all ones."""

    def __init__(self, savepath, zero_diag=False, device=None):
        super().__init__(savepath, zero_diag, device)

    def compute_all(self, graphs):
        return [torch.ones((_num_nodes(g), _num_nodes(g))) for g in graphs]


class LapEncoding(PositionEncoding):
    """:118-168 -- first ``dim`` non-trivial Laplacian eigenvectors (ascending eigenvalue), zero padded.
    The reference uses ``np.linalg.eig`` (unit-norm eigenvectors); eigenvector signs (and bases of
    repeated eigenvalues) are not unique -- the drivers randomise the sign anyway
    (run_transformer_gengcn_SBM_cv.py:159-164)."""

    def __init__(self, dim, use_edge_attr=False, normalization=None, device=None):
        super().__init__(None, False, device)
        self.pos_enc_dim, self.normalization, self.use_edge_attr = dim, normalization, use_edge_attr

    def compute_all(self, graphs):
        rw = self.normalization == 'rw'      # eigenvectors of L_rw = S^-1 (eigenvectors of L_sym), renormalised
        out = [None] * len(graphs)
        for idx in _size_sorted_chunks(graphs):
            L, ns, deg = dense_laplacians([graphs[i] for i in idx], 'sym' if rw else self.normalization,
                                          device=self.device, return_deg=True, weighted=self.use_edge_attr)
            sym = _is_symmetric(L).cpu().numpy()
            nmax = L.shape[1]
            # push the padding's zero eigenvalues to the top so real eigenpairs come first, ascending
            pad = (torch.arange(nmax, device=L.device).unsqueeze(0)
                   >= torch.from_numpy(ns).to(L.device).unsqueeze(1))
            w, U = torch.linalg.eigh(L + torch.diag_embed(pad.to(L.dtype) * 1e6))
            if rw:
                U = U / _rw_scale(deg).unsqueeze(2)
                U = U / torch.linalg.vector_norm(U, dim=1, keepdim=True)      # np.linalg.eig returns unit vectors
            for j, (i, n) in enumerate(zip(idx, ns)):
                if not sym[j]:      # directed edge list: the general eigendecomposition the reference calls (:137-139)
                    Lj = dense_laplacians([graphs[i]], self.normalization, weighted=self.use_edge_attr)[0][0].numpy()
                    val, vec = np.linalg.eig(Lj)
                    pe = torch.from_numpy(np.real(vec[:, val.argsort()])[:, 1:self.pos_enc_dim + 1])
                else:
                    pe = U[j, :n, 1:self.pos_enc_dim + 1]
                    pe = pe[:, :max(0, min(self.pos_enc_dim, n - 1))]
                full = torch.zeros((n, self.pos_enc_dim), dtype=torch.float32)
                full[:, :pe.shape[1]] = pe.to(torch.float32).cpu()
                out[i] = full
        return out

    def apply_to(self, dataset):
        graphs = [dataset[i] for i in range(len(dataset))]
        dataset.lap_pe_list = self.compute_all(graphs)
        return dataset


POSENCODINGS = {
    "diffusion": DiffusionEncoding,
    "pstep": PStepRWEncoding,
    "adj": AdjEncoding,
}
