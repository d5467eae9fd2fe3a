"""CPU restatement of ``ChebConvDynamic`` -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/transformer/ChebNetDynamic.py line by line, including
the per-node materialisation of the filter (``:148-149``) and the un-fused
``bmm`` + ``propagate`` chain, so that timing it is timing the reference's
algorithm.  PARITY UNPINNED: the reference holds no test for this operator.
"""
import math

import torch
from torch import nn

from . import pyg17


def cheb_norm(edge_index, num_nodes, edge_weight, normalization, lambda_max, dtype=None,
              batch=None):
    """``ChebConvDynamic.__norm__`` -- ChebNetDynamic.py:108-130."""
    edge_index, edge_weight = pyg17.remove_self_loops(edge_index, edge_weight)      # :113
    edge_index, edge_weight = pyg17.get_laplacian(edge_index, edge_weight,         # :115-117
                                                  normalization, dtype, num_nodes)
    if batch is not None and lambda_max.numel() > 1:                                 # :119-120
        lambda_max = lambda_max[batch[edge_index[0]]]
    edge_weight = (2.0 * edge_weight) / lambda_max                                   # :122
    edge_weight = edge_weight.masked_fill(edge_weight == float('inf'), 0)            # :123
    edge_index, edge_weight = pyg17.add_self_loops(edge_index, edge_weight,          # :125-127
                                                   fill_value=-1.0, num_nodes=num_nodes)
    return edge_index, edge_weight


def cheb_conv_dynamic(x, edge_index, filter_coeff, batch=None, lambda_max=None, bias=None,
                      weight=None, learn_only_filter_order_coeff=False, normalization='sym',
                      edge_weight=None):
    """``ChebConvDynamic.forward`` -- ChebNetDynamic.py:132-189.

    x [R, Fin]; edge_index [2, E] int64; filter_coeff [K, G, Fin, Fout]
    (default mode) or [K, G] (``learn_only_filter_order_coeff``); batch [R]
    (int or float, models.py:179-182 passes float).
    """
    if normalization != 'sym' and lambda_max is None:                                # :135-137
        raise ValueError('You need to pass `lambda_max` to `forward() in`'
                         'case the normalization is non-symmetric.')
    if lambda_max is None:                                                           # :139-140
        lambda_max = torch.tensor(2.0, dtype=x.dtype, device=x.device)
    if not isinstance(lambda_max, torch.Tensor):                                     # :141-143
        lambda_max = torch.tensor(lambda_max, dtype=x.dtype, device=x.device)

    if batch is not None:                                                            # :146-155
        _, repeat_indices = torch.unique(batch, sorted=True, return_counts=True)
        if not learn_only_filter_order_coeff:
            w = torch.repeat_interleave(filter_coeff, repeat_indices, dim=1)
        else:
            filter_coeff = torch.repeat_interleave(filter_coeff, repeat_indices, dim=1)
            w = weight
    else:
        # the reference leaves ``weight`` unbound here (NameError) -- there is no
        # un-batched mode in FeTA; restated as a single graph.
        raise NameError("ChebConvDynamic.forward: `weight` is unbound when batch is None "
                        "(ChebNetDynamic.py:146-166)")

    edge_index, norm = cheb_norm(edge_index, x.size(0), edge_weight, normalization,  # :157-160
                                 lambda_max, dtype=x.dtype, batch=batch)

    Tx_0 = x
    Tx_1 = x
    if learn_only_filter_order_coeff:                                                # :164-167
        out = torch.matmul(filter_coeff[0].unsqueeze(1) * Tx_0, w[0])
    else:
        out = torch.bmm(Tx_0.unsqueeze(1), w[0]).squeeze()
    if w.size(0) > 1:                                                                # :170-175
        Tx_1 = pyg17.propagate_add(edge_index, x, norm, x.size(0))
        if learn_only_filter_order_coeff:
            out = out + torch.matmul(filter_coeff[1].unsqueeze(1) * Tx_1, w[1])
        else:
            out = out + torch.bmm(Tx_1.unsqueeze(1), w[1]).squeeze()
    for k in range(2, w.size(0)):                                                    # :177-184
        Tx_2 = pyg17.propagate_add(edge_index, Tx_1, norm, x.size(0))
        Tx_2 = 2. * Tx_2 - Tx_0
        if learn_only_filter_order_coeff:
            out = out + torch.matmul(filter_coeff[k].unsqueeze(1) * Tx_2, w[k])
        else:
            out = out + torch.bmm(Tx_2.unsqueeze(1), w[k]).squeeze()
        Tx_0, Tx_1 = Tx_1, Tx_2
    if bias is not None:                                                             # :186-187
        out = out + bias
    return out


class OracleChebConvDynamic(nn.Module):
    """Module shell with the reference's parameter names (ChebNetDynamic.py:80-106)."""

    def __init__(self, in_channels, out_channels, K, normalization='sym', bias=True,
                 learn_only_filter_order_coeff=False):
        super().__init__()
        assert K > 0
        assert normalization in [None, 'sym', 'rw'], 'Invalid normalization'
        self.in_channels, self.out_channels, self.K = in_channels, out_channels, K
        self.normalization = normalization
        self.learn_only_filter_order_coeff = learn_only_filter_order_coeff
        if learn_only_filter_order_coeff:
            self.weight = nn.Parameter(torch.empty(K, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.learn_only_filter_order_coeff:                                       # glorot :20-23
            stdv = math.sqrt(6.0 / (self.weight.size(-2) + self.weight.size(-1)))
            self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.fill_(0)

    def forward(self, x, edge_index, filter_coeff, edge_weight=None, batch=None, lambda_max=None):
        return cheb_conv_dynamic(
            x, edge_index, filter_coeff, batch=batch, lambda_max=lambda_max, bias=self.bias,
            weight=getattr(self, 'weight', None),
            learn_only_filter_order_coeff=self.learn_only_filter_order_coeff,
            normalization=self.normalization, edge_weight=edge_weight)
