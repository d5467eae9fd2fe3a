"""CPU restatement of the reference's ``ARMAConvDynamic`` (transformer/ChebNetDynamic.py:201-358).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference holds no test or
golden vector for this operator and PyG 1.7 is not installable here; the restatement follows the
reference lines literally (per-node weight materialisation, ``_batch_multiply_coeff`` bmm, PyG-1.7
``gcn_norm(add_self_loops=False)`` and ``propagate``).
"""
import torch

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
