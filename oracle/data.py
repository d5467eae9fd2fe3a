"""CPU restatement of the collate / index builders and position encodings --
TEST INFRASTRUCTURE ONLY.

Follows /root/reference/transformer/data.py:113-225 (GraphDataset_v2), :229-344
(GraphDataset_ogb), :346-460 (GraphDataset_sbm) with their Python per-graph and
per-node loops, and transformer/position_encoding.py:55-72,118-161.  The integer
outputs (mask, edge_indices, batch_indices, feature_indices_to_gather) are fully
determined by those lines and are the bit-exact target of the product's
vectorised host collate and GPU batch builder.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import pyg17


class Graph(object):
    """Minimal stand-in for a torch_geometric ``Data`` object."""

    def __init__(self, x, edge_index, y, edge_attr=None):
        self.x, self.edge_index, self.y, self.edge_attr = x, edge_index, y, edge_attr
        self.num_nodes = x.shape[0]
        self.pe = None
        self.lap_pe = None
        self.degree = None
        self.x_onehot = None


def compute_degree(g):
    """data.py:142-146: ``1 / sqrt(1 + degree(edge_index[0]))``."""
    return 1. / torch.sqrt(1. + pyg17.degree(g.edge_index[0], g.num_nodes))


def one_hot(g, n_tags):
    """data.py:153-159."""
    return F.one_hot(g.x.view(-1).long(), n_tags)


def _collate_common(batch, n_tags, n_features, kind):
    batch = list(batch)
    max_len = max(len(g.x) for g in batch)                                           # :165
    if n_tags is None:
        padded_x = torch.zeros((len(batch), max_len, n_features))                    # :168
    else:
        padded_x = torch.zeros((len(batch), max_len, n_tags))                        # :171
    mask = torch.zeros((len(batch), max_len), dtype=bool)                            # :172
    labels = []
    pos_enc = None
    use_pe = getattr(batch[0], 'pe', None) is not None                               # :179
    if use_pe:
        pos_enc = torch.zeros((len(batch), max_len, max_len))                        # :182
    lap_pos_enc = None
    use_lap_pe = getattr(batch[0], 'lap_pe', None) is not None                       # :188
    if use_lap_pe:
        lap_pe_dim = batch[0].lap_pe.shape[-1]
        lap_pos_enc = torch.zeros((len(batch), max_len, lap_pe_dim))                 # :191
    degree = None
    use_degree = getattr(batch[0], 'degree', None) is not None                       # :194
    if use_degree:
        degree = torch.zeros((len(batch), max_len))                                  # :196
    feature_indices_to_gather, edge_indices, batch_indices, edge_attrs = [], [], [], []
    node_offset = 0
    for i, g in enumerate(batch):                                                    # :202-221
        labels.append(g.y)
        g_len = len(g.x)
        if n_tags is None:
            padded_x[i, :g_len, :] = g.x
        else:
            padded_x[i, :g_len, :] = g.x_onehot
        mask[i, g_len:] = True
        if use_pe:
            pos_enc[i, :g_len, :g_len] = g.pe
        if use_lap_pe:
            lap_pos_enc[i, :g_len, :g.lap_pe.shape[-1]] = g.lap_pe
        if use_degree:
            degree[i, :g_len] = g.degree
        feature_indices_to_gather.extend([[i, node_idx] for node_idx in range(g_len)])
        edge_indices.append(g.edge_index + node_offset)
        batch_indices.extend([i] * g_len)
        if kind == 'ogb':
            edge_attrs.append(g.edge_attr)                                           # :338
        node_offset += g_len
    edge_indices = torch.cat(edge_indices, dim=1)                                    # :223
    if kind == 'sbm':
        labels_out = torch.cat(labels, dim=0)                                        # :457
    else:
        labels_out = torch.stack([torch.as_tensor(l) for l in labels], dim=0)        # default_collate
    out = (padded_x, mask, pos_enc, lap_pos_enc, degree, labels_out, edge_indices,
           torch.tensor(batch_indices), torch.tensor(feature_indices_to_gather))
    if kind == 'ogb':
        out = out + (torch.cat(edge_attrs, dim=0),)                                  # :342
    return out


def collate_v2(batch, n_tags=None, n_features=None):
    """GraphDataset_v2.collate_fn -- data.py:161-225."""
    return _collate_common(batch, n_tags, n_features, 'v2')


def collate_sbm(batch, n_tags=None, n_features=None):
    """GraphDataset_sbm.collate_fn -- data.py:394-460 (labels concatenated per node)."""
    return _collate_common(batch, n_tags, n_features, 'sbm')


def collate_ogb(batch, n_tags=None, n_features=None):
    """GraphDataset_ogb.collate_fn -- data.py:277-344 (also returns edge_attr)."""
    return _collate_common(batch, n_tags, n_features, 'ogb')


def _dense_laplacian(edge_index, num_nodes, normalization, edge_weight=None):
    ei, ew = pyg17.get_laplacian(edge_index, edge_weight, normalization=normalization,
                                 dtype=torch.float32, num_nodes=num_nodes)
    L = np.zeros((num_nodes, num_nodes), dtype=np.float32)
    np.add.at(L, (ei[0].numpy(), ei[1].numpy()), ew.numpy())      # to_scipy_sparse_matrix sums dups
    return L


def diffusion_pe(edge_index, num_nodes, beta=1.0, normalization=None, edge_weight=None):
    """DiffusionEncoding.compute_pe -- position_encoding.py:65-72: ``expm(-beta L)``."""
    from scipy.linalg import expm
    L = _dense_laplacian(edge_index, num_nodes, normalization, edge_weight)   # use_edge_attr: :66, :81, :129
    return torch.from_numpy(expm(-beta * L))


def pstep_pe(edge_index, num_nodes, p=1, beta=0.5, normalization=None, edge_weight=None):
    """PStepRWEncoding.compute_pe -- position_encoding.py:83-93: ``(I - beta L)^p``."""
    L = _dense_laplacian(edge_index, num_nodes, normalization, edge_weight)   # use_edge_attr: :66, :81, :129
    M = np.eye(num_nodes, dtype=L.dtype) - beta * L
    tmp = M
    for _ in range(p - 1):
        tmp = tmp.dot(M)
    return torch.from_numpy(tmp)


def lap_pe(edge_index, num_nodes, dim, normalization=None, edge_weight=None):
    """LapEncoding.compute_pe -- position_encoding.py:127-161 (np.linalg.eig, ascending
    eigenvalues, drop the first eigenvector, zero-pad to ``dim`` columns)."""
    L = _dense_laplacian(edge_index, num_nodes, normalization, edge_weight)   # use_edge_attr: :66, :81, :129
    EigVal, EigVec = np.linalg.eig(L)
    idx = EigVal.argsort()
    EigVal, EigVec = EigVal[idx], np.real(EigVec[:, idx])
    eig_vec_pe = EigVec[:, 1:dim + 1]
    if eig_vec_pe.shape[1] < dim:
        pad = np.zeros((eig_vec_pe.shape[0], dim))
        pad[:, :eig_vec_pe.shape[1]] = eig_vec_pe
        eig_vec_pe = pad
    return torch.from_numpy(eig_vec_pe).float()
