"""Drop-in ``ChebConvDynamic`` (reference: transformer/ChebNetDynamic.py:29-198) and
``ARMAConvDynamic`` (:201-358).

Same constructor, ``forward`` signature, parameter names (``bias``; ``weight`` only in
``learn_only_filter_order_coeff`` mode) and output shape as the reference's PyG module, but the
whole forward is ONE fused sm_100a kernel (csrc/cheb.cu) behind the C ABI, with a matching fused
backward, over a CSR plan of the scaled Laplacian built once per mini-batch (csrc/cheb_plan.cu).
"""
import math
from collections import OrderedDict

import torch
from torch import nn

from . import ops


class ChebConvDynamic(nn.Module):
    def __init__(self, in_channels, out_channels, K, normalization='sym', bias=True,
                 learn_only_filter_order_coeff=False, **kwargs):
        # kwargs: MessagePassing options of the reference (aggr/flow/node_dim); only the defaults
        # the reference uses ('add', 'source_to_target', 0) are implemented
        aggr = kwargs.pop('aggr', 'add')
        flow = kwargs.pop('flow', 'source_to_target')
        node_dim = kwargs.pop('node_dim', 0)
        if aggr != 'add' or flow != 'source_to_target' or node_dim != 0 or kwargs:
            raise NotImplementedError("ChebConvDynamic(b200): only aggr='add', "
                                      "flow='source_to_target', node_dim=0 are implemented")
        super().__init__()
        assert K > 0
        assert normalization in [None, 'sym', 'rw'], 'Invalid normalization'
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.K = K
        self.normalization = normalization
        if learn_only_filter_order_coeff:
            self.weight = nn.Parameter(torch.empty(K, in_channels, out_channels))
        self.learn_only_filter_order_coeff = learn_only_filter_order_coeff
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self._plans = OrderedDict()      # small identity-keyed cache: one plan per mini-batch
        self.plan_hints = None           # set by a caller that knows max_nodes on the host
        self.reset_parameters()

    def reset_parameters(self):
        if self.learn_only_filter_order_coeff:          # glorot, ChebNetDynamic.py:20-23
            stdv = math.sqrt(6.0 / (self.weight.size(-2) + self.weight.size(-1)))
            self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.fill_(0)

    def __deepcopy__(self, memo):       # plans hold device buffers tied to one batch: never cloned
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, OrderedDict() if k == '_plans' else copy.deepcopy(v, memo))
        return new

    # -- plan management -------------------------------------------------------------------
    def get_plan(self, edge_index, batch, num_rows, num_graphs, lambda_max, hints=None):
        """The reference recomputes ``__norm__`` (:157-160) and ``unique`` (:148) on every call;
        here the CSR plan is cached per (edge_index, batch) identity + version."""
        key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape),
               None if batch is None else (batch.data_ptr(), batch._version, batch.dtype),
               num_rows, num_graphs, float(lambda_max), None if hints is None else tuple(sorted(hints.items())))
        hit = self._plans.get(key)
        if hit is not None:
            self._plans.move_to_end(key)
            return hit
        plan = ops.build_cheb_plan(edge_index, batch, num_rows, num_graphs, lambda_max, hints=hints,
                                   norm=getattr(self, '_plan_norm', ops.NORM_CHEB_SYM))
        plan._keep = (edge_index, batch)     # pins the storage so the identity key cannot be recycled
        self._plans[key] = plan
        while len(self._plans) > 4:
            self._plans.popitem(last=False)
        return plan

    def forward(self, x, edge_index, filter_coeff, edge_weight=None, batch=None, lambda_max=None,
                plan=None):
        """ChebNetDynamic.py:132-189.  ``plan`` (extension): a prebuilt ``ops.ChebPlan``."""
        if self.normalization != 'sym' and lambda_max is None:
            raise ValueError('You need to pass `lambda_max` to `forward() in`'
                             'case the normalization is non-symmetric.')
        if self.normalization != 'sym':
            raise NotImplementedError("ChebConvDynamic(b200): only normalization='sym' is implemented "
                                      "(the reference drivers never pass another, SURVEY.md section 5)")
        if edge_weight is not None:
            raise NotImplementedError("ChebConvDynamic(b200): edge_weight is not implemented "
                                      "(never passed by the reference, models.py:360)")
        if lambda_max is None:
            lambda_max = 2.0
        if isinstance(lambda_max, torch.Tensor):
            if lambda_max.numel() != 1:
                raise NotImplementedError("ChebConvDynamic(b200): per-graph lambda_max is not implemented")
            lambda_max = float(lambda_max)
        if float(lambda_max) != 2.0:
            # __norm__ (:115-127) leaves a diagonal of 2/lambda_max - 1 (the +1 loops of get_laplacian scaled,
            # then add_self_loops(-1)); it vanishes only at lambda_max = 2, which is all the plan stores.
            raise NotImplementedError("ChebConvDynamic(b200): lambda_max != 2.0 is not implemented (the reference "
                                      "drivers never pass it: models.py:360 -> None -> 2.0)")
        if batch is None:
            raise NameError("ChebConvDynamic.forward: `weight` is unbound when batch is None "
                            "(ChebNetDynamic.py:146-166) -- pass `batch`")
        if self.learn_only_filter_order_coeff:
            # out = sum_k (c_k[g] * T_k) W_k  ==  sum_k T_k (c_k[g] W_k): same fused kernel
            theta = filter_coeff.reshape(filter_coeff.shape[0], filter_coeff.shape[1], 1, 1) \
                * self.weight.unsqueeze(1)
        else:
            theta = filter_coeff
        if theta.dim() != 4:
            raise ValueError("filter_coeff must be [K, G, in, out]; got %s" % (tuple(theta.shape),))
        R = x.size(0)
        G = theta.size(1)
        if plan is None:
            plan = self.get_plan(edge_index, batch, R, G, lambda_max, hints=self.plan_hints)
        out = ops.cheb_filter(x, theta, self.bias, plan)
        return out.squeeze()        # the reference's bmm(...).squeeze(), :167

    def __repr__(self):
        return '{}({}, {}, K={}, normalization={})'.format(
            self.__class__.__name__, self.in_channels, self.out_channels, self.K, self.normalization)


class ARMAConvDynamic(nn.Module):
    """Drop-in for the reference's ``ARMAConvDynamic`` (transformer/ChebNetDynamic.py:201-358) as the
    encoder instantiates it (``num_layers=1``, transformer/models.py:139).

    Same constructor, parameter names and shapes (``init_weight [K,Fin,Fout]``, ``weight
    [max(1,T-1),K,Fout,Fout]``, ``root_weight [T,K,Fin,Fout]``, ``bias [T,K,1,Fout]``) so reference
    checkpoints load; ``forward(x, edge_index, filter_coeff [G, 2K], batch=...)`` is one fused kernel
    (csrc/arma.cu) over a ``gcn_norm`` CSR plan.  Configurations the fused kernel does not cover raise.
    """
    get_plan = ChebConvDynamic.get_plan
    __deepcopy__ = ChebConvDynamic.__deepcopy__

    def __init__(self, in_channels, out_channels, num_stacks=1, num_layers=1, shared_weights=False,
                 act='relu', dropout=0., bias=True, **kwargs):
        aggr = kwargs.pop('aggr', 'add')
        if aggr != 'add' or kwargs:
            raise NotImplementedError("ARMAConvDynamic(b200): only aggr='add' is implemented")
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("ARMAConvDynamic(b200): only num_layers=1 is implemented (the only value "
                                      "the reference uses, transformer/models.py:139)")
        if in_channels != out_channels:
            raise ValueError("ARMAConvDynamic: in_channels must equal out_channels "
                             "(_batch_multiply_coeff reshapes with shape[-2] twice, ChebNetDynamic.py:284)")
        if not (act == 'relu' or isinstance(act, nn.ReLU)):
            raise NotImplementedError("ARMAConvDynamic(b200): only the default ReLU activation is implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_stacks, self.num_layers = num_stacks, num_layers
        self.act = nn.ReLU()
        self.shared_weights = shared_weights
        self.dropout = dropout
        K, T, F_in, F_out = num_stacks, num_layers, in_channels, out_channels
        T = 1 if shared_weights else T
        self.init_weight = nn.Parameter(torch.empty(K, F_in, F_out))
        self.weight = nn.Parameter(torch.empty(max(1, T - 1), K, F_out, F_out))   # unused when num_layers == 1
        self.root_weight = nn.Parameter(torch.empty(T, K, F_in, F_out))
        if bias:
            self.bias = nn.Parameter(torch.empty(T, K, 1, F_out))
        else:
            self.register_parameter('bias', None)
        self._plans = OrderedDict()
        self._plan_norm = ops.NORM_GCN
        self.plan_hints = None
        self.reset_parameters()

    def reset_parameters(self):
        for w in (self.init_weight, self.weight, self.root_weight):    # glorot, ChebNetDynamic.py:268-272
            stdv = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
            w.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.fill_(0)

    def forward(self, x, edge_index, filter_coeff, edge_weight=None, batch=None, plan=None):
        """ChebNetDynamic.py:297-346.  ``plan`` (extension): a prebuilt NORM_GCN ``ops.ChebPlan``."""
        if edge_weight is not None:
            raise NotImplementedError("ARMAConvDynamic(b200): edge_weight is not implemented "
                                      "(never passed by the reference, models.py:363)")
        if batch is None:
            raise NotImplementedError("ARMAConvDynamic(b200): pass `batch` (per-node filter_coeff rows without "
                                      "a batch vector are not implemented)")
        if filter_coeff.dim() != 2 or filter_coeff.size(1) != 2 * self.num_stacks:
            raise ValueError("filter_coeff must be [G, 2*num_stacks]; got %s" % (tuple(filter_coeff.shape),))
        R, G = x.size(0), filter_coeff.size(0)
        if plan is None:
            plan = self.get_plan(edge_index, batch, R, G, 2.0, hints=self.plan_hints)
        x_root = None
        if self.training and self.dropout > 0:
            x_root = torch.nn.functional.dropout(x, p=self.dropout, training=True)    # :335
        bias = None if self.bias is None else self.bias[0].reshape(self.num_stacks, self.out_channels)
        return ops.arma_filter(x, filter_coeff, self.init_weight, self.root_weight[0], bias, plan, x_root=x_root)

    def __repr__(self):
        return '{}({}, {}, num_stacks={}, num_layers={})'.format(
            self.__class__.__name__, self.in_channels, self.out_channels, self.num_stacks, self.num_layers)
