"""Import shim that lets the UNMODIFIED reference Python run in this container.

TEST INFRASTRUCTURE ONLY -- used by ``make_golden_from_reference.py`` (and by the CPU tests that re-run the
reference when ``/root/reference`` is present).  Nothing here is on the product path.

What stands in the way of ``import transformer.models`` here:
  * ``torch_geometric`` (pinned ``torch-geometric=1.7``, README.md:24 of the reference), ``torch_scatter``,
    ``torch_sparse`` and ``ogb`` are not installed and cannot be (no network);
  * ``transformer/layers.py`` is a byte copy of ``gckn/layers.py``: the class ``models.py:4`` imports from it,
    ``DiffTransformerEncoderLayer``, exists nowhere in the tree;
  * ``np.long`` (``models.py:264``) is gone from NumPy 2, and ``utils.DEVICE`` is only set by ``init_device()``.

``install()`` registers minimal stand-ins for exactly the third-party symbols the hot path touches -- each one
a restatement of the published torch_geometric 1.7 / torch_scatter / ogb behaviour, written independently of
``oracle/pyg17.py`` (scatter via ``Tensor.scatter_add_`` here, ``index_add_`` there) so that agreement between
the reference run and the oracle is agreement between two restatements of the dependency plus the reference's
own, unmodified op sequence:

    ChebNetDynamic.py   ChebConvDynamic / ARMAConvDynamic .forward/.__norm__/.message  -- reference code
    models.py           DiffTransformerEncoderGenGCN (+ get_filter_coefficients, filter), the three heads,
                        GlobalAvg1D                                                   -- reference code
    data.py             GraphDataset_v2 / _sbm / _ogb  collate_fn                     -- reference code
    position_encoding.py  Diffusion / PStepRW / Adj / Full / Lap encodings (compute_pe)  -- reference code
    MessagePassing.propagate, get_laplacian, remove/add_self_loops, gcn_norm,
    GCNConv, global_mean_pool, degree, to_scipy_sparse_matrix, to_dense_adj,
    AtomEncoder, BondEncoder                                                         -- this shim
    DiffTransformerEncoderLayer                                                      -- oracle/layers.py (F1)
"""
import inspect
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("FETA_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "transformer", "ChebNetDynamic.py"))


# ------------------------------------------------------------------------------------------------
# torch_scatter / torch_geometric.utils  (published PyG-1.7 semantics)
# ------------------------------------------------------------------------------------------------
def _scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim in (0, -2) or src.dim() == 1
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = (dim_size,) + tuple(src.shape[1:])
    if out is None:
        out = src.new_zeros(shape)
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    return out.scatter_add_(0, idx, src)


def _maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return num_nodes
    return int(edge_index.max()) + 1 if edge_index.numel() else 0


def _remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def _add_self_loops(edge_index, edge_weight=None, fill_value=1., num_nodes=None):
    N = _maybe_num_nodes(edge_index, num_nodes)
    loops = torch.arange(N, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        edge_weight = torch.cat([edge_weight, edge_weight.new_full((N,), fill_value)], dim=0)
    return torch.cat([edge_index, loops], dim=1), edge_weight


def _add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1., num_nodes=None):
    N = _maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index
    off = row != col
    loops = torch.arange(N, dtype=row.dtype, device=row.device).unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        lw = edge_weight.new_full((N,), fill_value)
        on = ~off
        if int(on.sum()) > 0:
            lw[row[on]] = edge_weight[on]
        edge_weight = torch.cat([edge_weight[off], lw], dim=0)
    return torch.cat([edge_index[:, off], loops], dim=1), edge_weight


def _get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    assert normalization in (None, 'sym', 'rw')
    edge_index, edge_weight = _remove_self_loops(edge_index, edge_weight)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    N = _maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index
    deg = _scatter_add(edge_weight, row, 0, dim_size=N)
    if normalization is None:
        edge_index, _ = _add_self_loops(edge_index, num_nodes=N)
        edge_weight = torch.cat([-edge_weight, deg], dim=0)
    elif normalization == 'sym':
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float('inf'), 0)
        edge_weight = dis[row] * edge_weight * dis[col]
        edge_index, edge_weight = _add_self_loops(edge_index, -edge_weight, fill_value=1., num_nodes=N)
    else:
        di = 1.0 / deg
        di.masked_fill_(di == float('inf'), 0)
        edge_weight = di[row] * edge_weight
        edge_index, edge_weight = _add_self_loops(edge_index, -edge_weight, fill_value=1., num_nodes=N)
    return edge_index, edge_weight


def _to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    """``torch_geometric.utils.to_scipy_sparse_matrix`` (call sites position_encoding.py:69,85,133): COO matrix,
    ones when no attribute is given; duplicates are summed when the caller converts (``.tocsc()``)."""
    import scipy.sparse
    row, col = edge_index.cpu()
    if edge_attr is None:
        edge_attr = torch.ones(row.size(0))
    else:
        edge_attr = edge_attr.view(-1).cpu()
        assert edge_attr.size(0) == row.size(0)
    N = _maybe_num_nodes(edge_index, num_nodes)
    return scipy.sparse.coo_matrix((edge_attr.numpy(), (row.numpy(), col.numpy())), (N, N))


def _to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    """``torch_geometric.utils.to_dense_adj`` for a single graph (call site position_encoding.py:105): ``[1, N, N]``,
    duplicate edges summed."""
    assert batch is None and edge_attr is None
    N = max_num_nodes if max_num_nodes is not None else _maybe_num_nodes(edge_index)
    adj = torch.zeros(N * N)
    adj.scatter_add_(0, edge_index[0] * N + edge_index[1], torch.ones(edge_index.size(1)))
    return adj.view(1, N, N)


def _degree(index, num_nodes=None, dtype=None):
    N = _maybe_num_nodes(index, num_nodes)
    out = torch.zeros((N,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones((index.size(0),)))


def _gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True, dtype=None):
    fill = 2. if improved else 1.
    N = _maybe_num_nodes(edge_index, num_nodes)
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    if add_self_loops:
        edge_index, edge_weight = _add_remaining_self_loops(edge_index, edge_weight, fill, N)
    row, col = edge_index[0], edge_index[1]
    deg = _scatter_add(edge_weight, col, 0, dim_size=N)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float('inf'), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


def _global_mean_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    s = _scatter_add(x, batch, 0, dim_size=size)
    cnt = _scatter_add(torch.ones(batch.size(0), dtype=x.dtype, device=x.device), batch, 0, dim_size=size)
    return s / cnt.clamp(min=1).view(-1, 1)


def _global_max_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return torch.stack([x[batch == g].max(dim=0).values for g in range(size)], dim=0)


class MessagePassing(nn.Module):
    """The slice of PyG-1.7 ``MessagePassing`` the path uses: ``propagate`` gathers ``x_j = x[edge_index[0]]``
    (flow='source_to_target'), calls the SUBCLASS's ``message`` with the keyword arguments its signature names,
    and scatter-adds at ``edge_index[1]``."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2, **kwargs):
        super().__init__()
        assert aggr == 'add' and flow == 'source_to_target'
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim      # PyG-1.7 default node_dim = -2
        self.fuse = False

    def propagate(self, edge_index, size=None, **kwargs):
        assert torch.is_tensor(edge_index)
        x = kwargs.get('x')
        dim = self.node_dim
        n = x.size(dim) if size is None else size[1]
        src, dst = edge_index[0], edge_index[1]
        names = [p for p in inspect.signature(self.message).parameters]
        args = {}
        for name in names:
            if name.endswith('_j'):
                args[name] = kwargs[name[:-2]].index_select(dim, src)
            elif name.endswith('_i'):
                args[name] = kwargs[name[:-2]].index_select(dim, dst)
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        if msg.dim() == 2:
            return _scatter_add(msg, dst, 0, dim_size=n)
        shape = list(msg.shape)
        shape[dim] = n
        return msg.new_zeros(shape).index_add_(dim if dim >= 0 else msg.dim() + dim, dst, msg)

    def message(self, x_j):
        return x_j


class GCNConv(MessagePassing):
    """PyG-1.7 ``GCNConv``: ``weight [in, out]`` (glorot), ``bias`` (zeros); X W, normalised propagate, + bias."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=True,
                 normalize=True, bias=True, **kwargs):
        super().__init__(aggr='add')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.add_self_loops, self.normalize = improved, add_self_loops, normalize
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        a = (6.0 / (in_channels + out_channels)) ** 0.5
        self.weight.data.uniform_(-a, a)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = _gcn_norm(edge_index, edge_weight, x.size(0), self.improved,
                                                self.add_self_loops, dtype=x.dtype)
        x = torch.matmul(x, self.weight)
        out = self.propagate(edge_index, x=x, edge_weight=edge_weight, size=None)
        if self.bias is not None:
            out = out + self.bias
        return out

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the hot path (SURVEY.md section 2); the shim does not provide it")


class _Embeds(nn.Module):
    dims = ()
    attr = ''

    def __init__(self, emb_dim):
        super().__init__()
        lst = nn.ModuleList()
        for d in self.dims:
            e = nn.Embedding(d, emb_dim)
            nn.init.xavier_uniform_(e.weight.data)
            lst.append(e)
        setattr(self, self.attr, lst)

    def forward(self, x):
        out = 0
        lst = getattr(self, self.attr)
        for i in range(x.shape[1]):
            out = out + lst[i](x[:, i])
        return out


class AtomEncoder(_Embeds):
    """ogb.graphproppred.mol_encoder.AtomEncoder (published feature vocabulary sizes)."""
    dims = (119, 4, 12, 12, 10, 6, 6, 2, 2)
    attr = 'atom_embedding_list'


class BondEncoder(_Embeds):
    dims = (5, 6, 2)
    attr = 'bond_embedding_list'


class Data(object):
    """Minimal ``torch_geometric.data.Data`` for the reference's collates (data.py reads ``x``, ``edge_index``,
    ``y``, ``num_nodes``, ``edge_attr`` and attaches ``x_onehot`` / ``pe`` / ``lap_pe`` / ``degree``)."""

    def __init__(self, x, edge_index, y, edge_attr=None):
        self.x, self.edge_index, self.y, self.edge_attr = x, edge_index, y, edge_attr

    @property
    def num_nodes(self):
        return self.x.shape[0]


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_INSTALLED = {}


def install(layer_cls=None):
    """Register the stand-ins, import the reference's modules and return them as a namespace.

    ``layer_cls``: the class to expose as ``transformer.layers.DiffTransformerEncoderLayer`` (default: the
    oracle's restatement, the only definition there is -- SURVEY.md F1)."""
    if _INSTALLED:
        return _INSTALLED['ns']
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if not hasattr(np, 'long'):
        np.long = int                                   # models.py:264 (NumPy < 1.24 alias)
    Opt = type(None)
    _module('torch_geometric')
    _module('torch_geometric.typing', OptTensor=Opt, Adj=Opt, PairTensor=Opt, Size=Opt)
    utils = _module('torch_geometric.utils', remove_self_loops=_remove_self_loops, add_self_loops=_add_self_loops,
                    add_remaining_self_loops=_add_remaining_self_loops, get_laplacian=_get_laplacian,
                    degree=_degree, to_scipy_sparse_matrix=_to_scipy_sparse_matrix, to_dense_adj=_to_dense_adj)
    _module('torch_geometric.utils.num_nodes', maybe_num_nodes=_maybe_num_nodes)
    nn_mod = _module('torch_geometric.nn', global_mean_pool=_global_mean_pool, global_max_pool=_global_max_pool,
                     GCNConv=GCNConv, ChebConv=_Unavailable, MessagePassing=MessagePassing)
    conv = _module('torch_geometric.nn.conv', MessagePassing=MessagePassing, GCNConv=GCNConv)
    gcn_conv = _module('torch_geometric.nn.conv.gcn_conv', gcn_norm=_gcn_norm, GCNConv=GCNConv)
    tg = sys.modules['torch_geometric']
    tg.utils, tg.nn, tg.typing = utils, nn_mod, sys.modules['torch_geometric.typing']
    nn_mod.conv, conv.gcn_conv = conv, gcn_conv
    _module('torch_scatter', scatter_add=_scatter_add)

    class SparseTensor(object):                          # only named in type annotations / isinstance checks
        pass

    def _na(*a, **k):
        raise NotImplementedError("torch_sparse is not on the hot path")

    _module('torch_sparse', SparseTensor=SparseTensor, matmul=_na, fill_diag=_na, sum=_na, mul=_na, spspmm=_na)
    _module('ogb')
    _module('ogb.graphproppred', PygGraphPropPredDataset=object)
    _module('ogb.graphproppred.mol_encoder', AtomEncoder=AtomEncoder, BondEncoder=BondEncoder)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    pkg = importlib.import_module('transformer')         # the reference's package (__init__ is empty)
    if layer_cls is None:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        if root not in sys.path:
            sys.path.insert(0, root)
        from oracle.layers import OracleDiffTransformerEncoderLayer as layer_cls
    layers = _module('transformer.layers', DiffTransformerEncoderLayer=layer_cls)
    pkg.layers = layers
    ref_utils = importlib.import_module('transformer.utils')
    ref_utils.DEVICE = 'cpu'                             # what init_device() sets without a GPU (utils.py:3-5)
    ns = types.SimpleNamespace(
        utils=ref_utils,
        cheb=importlib.import_module('transformer.ChebNetDynamic'),
        models=importlib.import_module('transformer.models'),
        data=importlib.import_module('transformer.data'),
        pe=importlib.import_module('transformer.position_encoding'),
        Data=Data)
    _INSTALLED['ns'] = ns
    return ns
