"""CPU restatement of the reference's ``ARMAConvDynamic`` (transformer/ChebNetDynamic.py:201-358).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference holds no test or
golden vector for this operator and PyG 1.7 is not installable here; the restatement follows the
reference lines literally (per-node weight materialisation, ``_batch_multiply_coeff`` bmm, PyG-1.7
``gcn_norm(add_self_loops=False)`` and ``propagate``).
"""
import math

import torch
from torch import nn

from .pyg17 import gcn_norm, propagate_add


def batch_multiply_coeff(x, w):
    """``ARMAConvDynamic._batch_multiply_coeff`` (ChebNetDynamic.py:274-295), shape for shape.

    x ``[1, R, F]``, w ``[R, K, F, F]`` -> ``[K, R, F]``.
    """
    x1 = x.permute([1, 0, 2]).unsqueeze(2)                       # :279
    x1 = x1.repeat([1, w.shape[1], 1, 1])                        # :280
    x11 = x1.reshape([-1, 1, x.shape[-1]])                       # :283
    w1 = w.reshape([-1, w.shape[-2], w.shape[-2]])               # :284 (in == out)
    x2 = torch.bmm(x11, w1)                                      # :285
    x21 = x2.reshape([-1, w.shape[1], 1, w.shape[-1]])           # :288
    x22 = x21.permute([1, 0, 2, 3])                              # :289
    return x22.reshape([w.shape[1], x.shape[1], x2.shape[-1]])   # :290


def arma_conv_dynamic(x, edge_index, filter_coeff, batch, init_weight, weight, root_weight, bias,
                      num_stacks, num_layers=1, shared_weights=False, x_root=None):
    """``ARMAConvDynamic.forward`` (ChebNetDynamic.py:297-346) with ``act = ReLU``.

    ``x_root`` stands in for ``F.dropout(x)`` of :335 (None = no dropout).
    """
    n = x.size(0)
    edge_index, edge_weight = gcn_norm(edge_index, None, n, add_loops=False, dtype=x.dtype)   # :302-304
    _, counts = torch.unique(batch, sorted=True, return_counts=True)                          # :313
    fc = torch.repeat_interleave(filter_coeff, counts, dim=0)                                 # :314
    fa = fc[:, :num_stacks].unsqueeze(-1).unsqueeze(-1)                                       # :315
    fb = fc[:, num_stacks:].unsqueeze(-1).unsqueeze(-1)                                       # :316
    x = x.unsqueeze(-3)                                                                       # :318
    root_in = x if x_root is None else x_root.unsqueeze(-3)
    out = x
    for t in range(num_layers):
        if t == 0:
            out = batch_multiply_coeff(out, init_weight * fa)                                 # :323-324
        else:
            raise NotImplementedError("num_layers > 1: the reference's _batch_multiply_coeff cannot take the "
                                      "[K, R, F] tensor of the second layer (:327-328)")
        out = torch.stack([propagate_add(edge_index, out[k], edge_weight, n)                  # :332-333
                           for k in range(out.shape[0])], dim=0)
        w = root_weight[0 if shared_weights else t] * fb                                      # :337
        out = out + batch_multiply_coeff(root_in, w)                                          # :338
        if bias is not None:
            out = out + bias[0 if shared_weights else t]                                      # :340-341
        out = torch.relu(out)                                                                 # :343-344
    return out.mean(dim=-3)                                                                   # :346


class OracleARMAConvDynamic(nn.Module):
    """Module shell with the reference's parameter names and shapes (ChebNetDynamic.py:244-272)."""

    def __init__(self, in_channels, out_channels, num_stacks=1, num_layers=1, shared_weights=False,
                 dropout=0., bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_stacks, self.num_layers = num_stacks, num_layers
        self.shared_weights, self.dropout = shared_weights, dropout
        K, T, F_in, F_out = num_stacks, num_layers, in_channels, out_channels
        T = 1 if shared_weights else T
        self.init_weight = nn.Parameter(torch.empty(K, F_in, F_out))                # :257
        self.weight = nn.Parameter(torch.empty(max(1, T - 1), K, F_out, F_out))     # :258
        self.root_weight = nn.Parameter(torch.empty(T, K, F_in, F_out))             # :259
        if bias:
            self.bias = nn.Parameter(torch.empty(T, K, 1, F_out))                   # :262
        else:
            self.register_parameter('bias', None)
        for w in (self.init_weight, self.weight, self.root_weight):                 # glorot, :269-271
            stdv = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
            w.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.fill_(0)                                                 # :272

    def forward(self, x, edge_index, filter_coeff, edge_weight=None, batch=None):
        assert edge_weight is None
        return arma_conv_dynamic(x, edge_index, filter_coeff, batch, self.init_weight, self.weight,
                                 self.root_weight, self.bias, self.num_stacks, self.num_layers,
                                 self.shared_weights)
