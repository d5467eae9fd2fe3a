"""Restatement of the torch_geometric==1.7 utilities the FeTA hot path calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  torch_geometric is a
third-party dependency of the reference that is neither vendored under
/root/reference nor installed here; the reference pins it only as
``torch-geometric=1.7`` in README.md:24.  Each function below restates the
published PyG-1.7 algorithm and names the reference call site that uses it.
All of it is plain index arithmetic plus ``index_add_``.
"""
import torch


def scatter_add(src, index, dim_size):
    """torch_scatter.scatter_add(src, index, dim=0, dim_size=...).

    Call sites: PyG ``get_laplacian`` degree (via ChebNetDynamic.py:115),
    ``gcn_norm`` degree (via models.py:282), ``MessagePassing`` aggregation
    (ChebNetDynamic.py:171,178).
    """
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def remove_self_loops(edge_index, edge_attr=None):
    """PyG ``remove_self_loops`` -- call site ChebNetDynamic.py:113."""
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[mask]


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    """PyG-1.7 ``add_self_loops`` -- call site ChebNetDynamic.py:125-127.

    Appends one (i, i) entry per node *unconditionally* (existing loops are not
    inspected) with weight ``fill_value``.
    """
    n = maybe_num_nodes(edge_index, num_nodes)
    loop_index = torch.arange(0, n, dtype=torch.long, device=edge_index.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        loop_weight = edge_weight.new_full((n,), fill_value)
        edge_weight = torch.cat([edge_weight, loop_weight], dim=0)
    edge_index = torch.cat([edge_index, loop_index], dim=1)
    return edge_index, edge_weight


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    """PyG-1.7 ``add_remaining_self_loops`` -- used by ``gcn_norm`` (models.py:282).

    Non-loop edges are kept, then one loop per node is appended whose weight is
    ``fill_value`` unless the input already held an (i, i) edge, in which case
    that edge's weight is kept as the loop weight.
    """
    n = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_index = torch.arange(0, n, dtype=row.dtype, device=row.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    new_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    if edge_weight is not None:
        inv_mask = ~mask
        loop_weight = torch.full((n,), fill_value, dtype=edge_weight.dtype,
                                 device=edge_weight.device)
        remaining = edge_weight[inv_mask]
        if remaining.numel() > 0:
            loop_weight[row[inv_mask]] = remaining
        edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
    return new_index, edge_weight


def get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    """PyG-1.7 ``get_laplacian`` -- call sites ChebNetDynamic.py:115,
    position_encoding.py:67,82,130.

    Degree is accumulated over ``row`` (= edge_index[0], the source).
    """
    assert normalization in (None, 'sym', 'rw')
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    n = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, row, n)
    if normalization is None:
        edge_index, _ = add_self_loops(edge_index, num_nodes=n)
        edge_weight = torch.cat([-edge_weight, deg], dim=0)
    elif normalization == 'sym':
        deg_inv_sqrt = deg.pow(-0.5)
        deg_inv_sqrt = deg_inv_sqrt.masked_fill(deg_inv_sqrt == float('inf'), 0)
        edge_weight = deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, fill_value=1.0,
                                                 num_nodes=n)
    else:
        deg_inv = 1.0 / deg
        deg_inv = deg_inv.masked_fill(deg_inv == float('inf'), 0)
        edge_weight = deg_inv[row] * edge_weight
        edge_index, edge_weight = add_self_loops(edge_index, -edge_weight, fill_value=1.0,
                                                 num_nodes=n)
    return edge_index, edge_weight


def propagate_add(edge_index, x, norm, num_nodes=None):
    """``MessagePassing.propagate(aggr='add', flow='source_to_target')`` with the
    message of ChebNetDynamic.py:192-193: ``out[col[e]] += norm[e] * x[row[e]]``.
    """
    n = x.size(0) if num_nodes is None else num_nodes
    row, col = edge_index[0], edge_index[1]
    msg = norm.view(-1, 1) * x.index_select(0, row)
    return scatter_add(msg, col, n)


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False,
             add_loops=True, dtype=None):
    """PyG-1.7 ``gcn_norm`` (dense-tensor branch) -- used by GCNConv (models.py:144,282).

    Degree is accumulated over ``col`` (the target).
    """
    fill_value = 2.0 if improved else 1.0
    n = maybe_num_nodes(edge_index, num_nodes)
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    if add_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight,
                                                           fill_value, n)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, col, n)
    deg_inv_sqrt = deg.pow(-0.5)
    deg_inv_sqrt = deg_inv_sqrt.masked_fill(deg_inv_sqrt == float('inf'), 0)
    return edge_index, deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]


def gcn_conv(x, edge_index, edge_weight, weight, bias):
    """PyG-1.7 ``GCNConv.forward`` -- call site models.py:282.

    ``weight`` has PyG-1.7's shape ``[in_channels, out_channels]`` (state_dict
    key ``gcn.weight``); ``x @ weight`` then normalised propagate then ``+ bias``.
    """
    edge_index, norm = gcn_norm(edge_index, edge_weight, x.size(0), False, True, dtype=x.dtype)
    x = torch.matmul(x, weight)
    out = propagate_add(edge_index, x, norm, x.size(0))
    if bias is not None:
        out = out + bias
    return out


def global_mean_pool(x, batch, size=None):
    """PyG-1.7 ``global_mean_pool`` (scatter mean, count clamped to >= 1) --
    call site models.py:283."""
    size = int(batch.max().item() + 1) if size is None else size
    out = scatter_add(x, batch, size)
    count = scatter_add(torch.ones(batch.size(0), dtype=x.dtype, device=x.device), batch, size)
    return out / count.clamp(min=1).view(-1, 1)


def degree(index, num_nodes=None, dtype=None):
    """PyG ``utils.degree`` -- call site data.py:145."""
    n = maybe_num_nodes(index, num_nodes)
    out = torch.zeros((n,), dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.index_add_(0, index, out.new_ones((index.size(0),)))
