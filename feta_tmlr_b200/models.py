"""Drop-in FeTA encoder + the model heads of the BASELINE configs
(reference: transformer/models.py:103-368, :487-551, :586-595, :598-725, :1008-1076).

Same class names, constructor/forward signatures, return arity and state_dict keys as the
reference.  What changed is *how* the filter path is computed:
  * get_filter_coefficients (models.py:240-287): no host loop, no ``.cpu()``, no all-pairs edge
    list, no [H*sum(n^2), ncoef] message tensor -- the all-ones GCN collapses to a per-node scalar
    computed by one kernel from the attention tile (csrc/coeff.cu), tanh + mean-pool in a second;
  * head stacking + packed gather (:177-185, :347) and scatter-back (:200-202): one kernel each;
  * ChebConvDynamic: one fused kernel (csrc/cheb.cu) over a CSR plan built once per batch.
The reference's un-tiled ``edge_index`` (only head 0's rows carry edges) is kept
by default (``tile_edges_per_head=False``) so numbers match the reference.
"""
import math

import torch
from torch import nn

from . import ops
from .ChebNetDynamic import ARMAConvDynamic, ChebConvDynamic
from .layers import DiffTransformerEncoderLayer, Linear

import os as _os
# forward_static builds the batch's padded-domain context + Laplacian plan on the side stream, concurrently with the
# encoder layers (FETA_STATIC_CONTEXT_SIDE_STREAM=0: on the main stream, where the first filtering layer needs them)
STATIC_CONTEXT_SIDE_STREAM = _os.environ.get("FETA_STATIC_CONTEXT_SIDE_STREAM", "1") == "1"
# ... and issues the filter-coefficient branch on its own stream so that its backward (parameter gradients only) runs
# beside the layers' backward chain (FETA_COEFF_BRANCH_STREAM=0: on the main stream)
COEFF_BRANCH_STREAM = _os.environ.get("FETA_COEFF_BRANCH_STREAM", "1") == "1"
import contextlib


# Every Linear of the heads / encoder glue is ``layers.Linear``: an ``nn.Linear`` (same parameters and state_dict keys)
# whose weight / bias gradients are reduced over the token axis by csrc/dense.cu -- the library's SIMT sgemm needs
# ~40 us for the [1024 x 256]^T [1024 x 256] weight gradient of ``encoder.linear`` alone (profiles/r2_launches_zinc.md).


class GCNConv(nn.Module):
    """Parameter shell with PyG-1.7 ``GCNConv`` names/shapes (``weight [in, out]``, ``bias``) so the
    reference's checkpoints load (models.py:144, :508).  On the hot path only ``weight.sum(0)`` and
    ``bias`` are used (the all-ones collapse); a general GCNConv forward is
    outside the path."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        stdv = math.sqrt(6.0 / (in_channels + out_channels))      # PyG glorot
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("GCNConv(b200) is a parameter shell: a general GCNConv forward is "
                                  "outside the hot path (SURVEY.md section 8)")


class BatchContext(object):
    """Index structures of one mini-batch shared by every filtering layer of a forward pass."""
    __slots__ = ("N", "B", "nmax", "H", "node_ptr", "batch_all_heads", "plan", "edge_index", "seg_lo", "seg_hi",
                 "real")


class DiffTransformerEncoderGenGCN(nn.TransformerEncoder):
    def __init__(self, d_model, num_heads, *args, num_coefficients=4, laplacian_norm='sym',
                 gnn_type='ChebConvDynamic', last_layer_filter=True,
                 learn_only_filter_order_coeff=False, use_skip_conn=True, tile_edges_per_head=False,
                 **kwargs):
        kwargs.setdefault('enable_nested_tensor', False)
        super().__init__(*args, **kwargs)
        if gnn_type not in ('ChebConvDynamic', 'ARMAConvDynamic'):
            raise NotImplementedError("gnn_type=%r: only 'ChebConvDynamic' and 'ARMAConvDynamic' are on the "
                                      "B200 hot path (SURVEY.md section 2)" % gnn_type)
        self.num_coefficients = num_coefficients
        dh = d_model // num_heads
        self.order = self.num_coefficients                                          # models.py:127,130
        if gnn_type == 'ARMAConvDynamic':                                           # :135-139
            self.num_coefficients = self.num_coefficients * 2
            self.spectral_gnns = ARMAConvDynamic(dh, dh, num_stacks=self.order, num_layers=1)
        elif learn_only_filter_order_coeff:
            self.spectral_gnns = ChebConvDynamic(dh, dh, self.num_coefficients,
                                                 normalization=laplacian_norm,
                                                 learn_only_filter_order_coeff=True)
        else:
            self.filter_in_channels = dh
            self.filter_out_channels = dh
            self.num_coefficients = self.order * dh * dh                            # :133
            self.spectral_gnns = ChebConvDynamic(dh, dh, self.order, normalization=laplacian_norm,
                                                 learn_only_filter_order_coeff=False)
        self.gcn = GCNConv(self.num_coefficients, self.num_coefficients)            # :144
        self.linear = Linear(self.num_coefficients, self.num_coefficients)       # :145
        self.linear_cat = Linear(2 * d_model, d_model)                           # :146
        self.gnn_type = gnn_type
        self.num_heads = num_heads
        self.last_layer_filter = last_layer_filter
        self.learn_only_filter_order_coeff = learn_only_filter_order_coeff
        self.use_skip_conn = use_skip_conn
        self.tile_edges_per_head = tile_edges_per_head

    # ---------------------------------------------------------------------------------------
    def batch_context(self, edge_index, feature_indices, batch, masks, nmax):
        """Per-mini-batch index structures (replaces models.py:180-185 and the host loop of
        :246-264).  No device->host synchronisation: graph count and the largest-graph bound
        come from the padded shapes."""
        ctx = BatchContext()
        B = masks.shape[0]
        H = self.num_heads
        N = feature_indices.shape[0]
        dev = masks.device
        ctx.N, ctx.B, ctx.nmax, ctx.H = N, B, nmax, H
        lens = (~masks).sum(dim=1)
        node_ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        node_ptr[1:] = torch.cumsum(lens, dim=0)
        ctx.node_ptr = node_ptr
        heads = torch.arange(H, device=dev, dtype=torch.int64).view(H, 1)
        ctx.batch_all_heads = (batch.to(torch.int64).view(1, N) + heads * B).reshape(-1)   # :181-182
        if self.tile_edges_per_head:
            ei = (edge_index.view(2, 1, -1) + (heads * N).view(1, H, 1)).reshape(2, -1)
        else:
            ei = edge_index                                                         # :186 (F4: un-tiled)
        ctx.edge_index = ei
        ctx.plan = self.spectral_gnns.get_plan(ei, ctx.batch_all_heads, H * N, H * B, 2.0,
                                               hints={'max_nodes': int(nmax), 'block_diagonal': True})
        return ctx

    # models.py:240-287
    def get_filter_coefficients(self, attn_weights, edge_index, feature_indices, batch, masks, ctx=None):
        if ctx is None:
            ctx = self.batch_context(edge_index, feature_indices, batch, masks, attn_weights.shape[2])
        s = ops.coeff_scalar(attn_weights, masks, ctx.node_ptr, ctx.N)             # :252-282, collapsed
        wbar = ops.colsum(self.gcn.weight)                                           # ones @ W
        pooled = ops.coeff_pool(s, ctx.plan.graph_ptr, wbar, self.gcn.bias)         # tanh + gap, :282-283
        pooled_coeff = self.linear(pooled)                                          # :284
        return pooled_coeff.reshape((self.num_heads, attn_weights.shape[0], pooled_coeff.shape[-1]))

    # models.py:346-368 (reference signature; the forward below uses the fused pack instead of :347)
    def filter(self, filter_coeff, graph_signal, edge_index, feature_indices, batch, spectral_gnn,
               eigenvalues=None, x=None, plan=None):
        if x is None:
            x = ops.gather_rows(graph_signal, feature_indices)                      # :347
        if self.gnn_type == 'ARMAConvDynamic':
            filter_coeff = filter_coeff.reshape((-1, self.order * 2))               # :361-363
            return spectral_gnn(x, edge_index, filter_coeff, batch=batch, plan=plan)
        if not self.learn_only_filter_order_coeff:
            filter_coeff = filter_coeff.reshape((-1, self.order, self.filter_in_channels,
                                                 self.filter_out_channels)).permute([1, 0, 2, 3])   # :357
        else:
            filter_coeff = filter_coeff.reshape((-1, self.order)).permute([1, 0])   # :359
        return spectral_gnn(x, edge_index, filter_coeff, batch=batch, plan=plan)    # :360

    # models.py:155-238
    def forward(self, src, pe, edge_index, feature_indices, batch, degree=None, mask=None,
                src_key_padding_mask=None, eigenvalues=None):
        output = src
        nmax, B, _ = src.shape
        H = self.num_heads
        if src_key_padding_mask is None:
            src_key_padding_mask = torch.zeros((B, nmax), dtype=torch.bool, device=src.device)
        coefficients = []
        allout_filtered = None
        num_layers = len(self.layers)
        ctx = None
        attn = None
        rowscale = None if degree is None else degree.transpose(0, 1).contiguous()   # once, not per layer
        for layer_num, mod in enumerate(self.layers):
            # the attention matrix is read only where the coefficients are built from it (:169-173) and, for the
            # last layer, returned (:238): the other layers never materialise it
            last = layer_num + 1 == num_layers
            output, attn, out_each_head = mod(output, pe=pe, degree=degree, src_mask=mask,
                                              src_key_padding_mask=src_key_padding_mask,
                                              need_heads=True, rowscale=rowscale,
                                              need_attn=last or not self.last_layer_filter)
            if self.last_layer_filter and layer_num + 1 != num_layers:              # :169-171
                continue
            if ctx is None:
                ctx = self.batch_context(edge_index, feature_indices, batch, src_key_padding_mask, nmax)
            coeff_all_heads = self.get_filter_coefficients(attn, edge_index, feature_indices, batch,
                                                           src_key_padding_mask, ctx=ctx)   # :173
            coeff = coeff_all_heads.reshape((H * B, coeff_all_heads.shape[2]))              # :178
            x = ops.pack_heads(out_each_head, feature_indices)                      # :179-185 + :347
            filtered = self.filter(coeff, None, ctx.edge_index, None, ctx.batch_all_heads,
                                   self.spectral_gnns, x=x, plan=ctx.plan)          # :186
            coefficients.append(coeff_all_heads)                                    # :198
            out_filtered = ops.unpack_heads(filtered.reshape(H * ctx.N, -1), feature_indices, B, nmax, H)  # :200-202
            if self.use_skip_conn:                                                  # :209-216
                allout_filtered = out_filtered if allout_filtered is None \
                    else allout_filtered + out_filtered
            else:
                allout_filtered = out_filtered
                output = allout_filtered
        if self.use_skip_conn:                                                      # :221-233
            if allout_filtered is not None:
                output = self.linear_cat(torch.cat((output, allout_filtered), dim=-1))
        else:
            if allout_filtered is not None:
                output = allout_filtered
        if self.norm is not None:
            output = self.norm(output)
        if coefficients:
            coeffs = torch.cat(coefficients, dim=0)
        else:
            coeffs = torch.empty((0, B, self.num_coefficients), device=src.device)
        return output, attn, coeffs.permute([1, 0, 2])                              # :238


    # ---------------------------------------------------------------------------------------
    # Static-shape ("padded domain") variant of forward(): every tensor shape depends only on
    # (B, Nmax, E_cap), never on the number of real nodes, so a whole training step can be captured
    # in ONE CUDA graph and replayed on new mini-batches (engine.GraphedTrainStep).  Every padded
    # slot (b, i) is treated as a node of graph b: padded slots carry x = 0, no edges, and are
    # masked out of the coefficient pooling and of the scattered-back result, so real nodes get
    # bit-for-bit the same arithmetic as forward().  `edge_index` holds the reference's packed
    # node ids, padded to a fixed width with (-1, -1) columns, which the plan builder ignores.
    def static_context(self, edge_index, masks, nmax):
        ctx = BatchContext()
        B, H = masks.shape[0], self.num_heads
        dev = masks.device
        ctx.N, ctx.B, ctx.nmax, ctx.H = B * nmax, B, nmax, H
        # graph sizes, their prefix, packed node id -> padded slot id (b * nmax + i) of every edge endpoint ((-1, -1)
        # padding columns stay -1; int32 edge lists -- the static batches' wire format -- are widened here), the
        # pooling segments and the real-row weights: two launches (csrc/collate.cu, feta_static_context)
        ei, ctx.node_ptr, ctx.seg_lo, ctx.seg_hi, ctx.real = ops.static_context_tensors(
            masks, edge_index, nmax, H, tile_heads=self.tile_edges_per_head)
        ctx.edge_index = ei
        key = (H, B, nmax, dev)
        cache = self.__dict__.setdefault('_static_batch_index', {})            # depends on the shapes only
        if key not in cache:
            cache.clear()
            cache[key] = torch.arange(H * B * nmax, device=dev, dtype=torch.int64) // nmax
        ctx.batch_all_heads = cache[key]
        ctx.plan = ops.build_cheb_plan(ei, ctx.batch_all_heads, H * B * nmax, H * B, 2.0,
                                       hints={'max_nodes': int(nmax), 'block_diagonal': True},
                                       norm=getattr(self.spectral_gnns, '_plan_norm', ops.NORM_CHEB_SYM))
        # kept so that engine.GraphedTrainStep.plan_guard_tripped() can read the guard word of the plans its
        # captured graphs rebuild in place on every replay
        sp = self.__dict__.setdefault('_static_plans', [])
        sp.append(ctx.plan)
        del sp[:-4]
        return ctx

    def forward_static(self, src, pe, edge_index, degree, masks):
        bn_rows = None
        if any(getattr(mod, 'batch_norm', False) for mod in self.layers):
            # the reference's BatchNorm1d runs over rows padded to the BATCH maximum (the layer flattens
            # [Nmax, B, d]); the static layout pads to a dataset-wide cap.  Rows i >= batch maximum are excluded
            # from the statistics by a 0/1 row weight (computed on the device: no synchronisation), which
            # reproduces the reference's statistics exactly; eval mode / other widths are not supported here.
            if not self.training or not ops.batchnorm_supported(src.shape[-1]):
                raise NotImplementedError("forward_static with batch_norm=True needs training mode and a d_model "
                                          "that divides 256 (fused BatchNorm with row weights)")
            batch_max = (~masks).sum(dim=1).max()
            bn_rows = (torch.arange(src.shape[0], device=src.device).view(-1, 1) < batch_max).to(torch.float32) \
                .expand(src.shape[0], src.shape[1]).reshape(-1).contiguous()
        output = src
        nmax, B, d = src.shape
        H = self.num_heads
        coefficients = []
        allout_filtered = None
        num_layers = len(self.layers)
        ctx = None
        attn = None
        rowscale = None if degree is None else degree.transpose(0, 1).contiguous()
        # The padded-domain context and the Laplacian plan depend on the batch's mask and edge list only: they are
        # built on the side stream while the encoder layers run (a parallel branch of the captured graph, ~15 launches
        # off the critical chain -- at the PATTERN shape the plan of 449k edges alone is ~100 us) and joined where the
        # first filtering layer needs them.  Every later use is ordered behind that join.
        side = None
        if src.is_cuda and STATIC_CONTEXT_SIDE_STREAM and num_layers > 0:
            main = torch.cuda.current_stream(src.device)
            side = ops._side_stream(src.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ctx = self.static_context(edge_index, masks, nmax)
        dev = src.device
        use_branch = bool(src.is_cuda and COEFF_BRANCH_STREAM)
        for layer_num, mod in enumerate(self.layers):
            # layers that feed the coefficients hand out a LazyAttention (ops.py): the coefficient scalar is
            # recomputed from q / k, so no layer of the static step materialises its attention matrix; the matrix
            # this method returns is therefore a LazyAttention too (``.materialize()`` for whoever wants to look)
            last = layer_num + 1 == num_layers
            filters = last or not self.last_layer_filter
            got = {}

            def filter_branch(attn_l, out_each_head):
                # The filter branch -- attention -> coefficient scalar -> tanh/mean pool -> Linear -> Chebyshev filter of
                # the stacked heads -- needs the attention outputs only, not the layer's token-wise tail (out_proj,
                # FFN, norms), and the two meet again at `linear_cat` / the skip sum.  It is issued on its own stream
                # right after the attention core: a parallel branch of the captured graph in the forward pass, and --
                # autograd runs a node's backward on the stream its forward ran on -- in the backward pass too.
                nonlocal side, ctx
                main = torch.cuda.current_stream(dev) if src.is_cuda else None
                if side is not None:                     # the context / plan branch joins here
                    main.wait_stream(side)
                    side = None
                if ctx is None:
                    ctx = self.static_context(edge_index, masks, nmax)
                branch = ops._side_stream(dev, 1) if use_branch else None
                if branch is not None:
                    branch.wait_stream(main)
                with (torch.cuda.stream(branch) if branch is not None else contextlib.nullcontext()):
                    s = ops.coeff_scalar(attn_l, masks, ctx.node_ptr, B * nmax, zero_fill=True)
                    pooled = ops.coeff_pool(s, ctx.seg_lo, ops.colsum(self.gcn.weight), self.gcn.bias,
                                            seg_hi=ctx.seg_hi)
                    coeff_all_heads = self.linear(pooled).reshape((H, B, -1))
                    coeff = coeff_all_heads.reshape((H * B, coeff_all_heads.shape[2]))
                    x = out_each_head.permute(2, 0, 1, 3).reshape(H * B * nmax, -1)      # padded-domain stacking
                    filtered = self.filter(coeff, None, ctx.edge_index, None, ctx.batch_all_heads,
                                           self.spectral_gnns, x=x, plan=ctx.plan)
                    got['out'] = filtered.reshape(H, B, nmax, -1).permute(2, 1, 0, 3).reshape(nmax, B, d) * ctx.real
                got['coeff'] = coeff_all_heads
                got['branch'] = branch

            output, attn, out_each_head = mod(output, pe=pe, degree=degree, src_key_padding_mask=masks,
                                              need_heads=True, rowscale=rowscale, bn_rows=bn_rows,
                                              need_attn='coeff' if filters else False,
                                              after_attention=filter_branch if filters else None)
            if not filters:
                continue
            if got['branch'] is not None:
                torch.cuda.current_stream(dev).wait_stream(got['branch'])
            coefficients.append(got['coeff'])
            out_filtered = got['out']
            if self.use_skip_conn:
                allout_filtered = out_filtered if allout_filtered is None else allout_filtered + out_filtered
            else:
                allout_filtered = out_filtered
                output = allout_filtered
        if side is not None:                         # (no layer filtered: the branch still has to rejoin)
            torch.cuda.current_stream(src.device).wait_stream(side)
        if self.use_skip_conn:
            if allout_filtered is not None:
                output = self.linear_cat(torch.cat((output, allout_filtered), dim=-1))
        else:
            if allout_filtered is not None:
                output = allout_filtered
        if self.norm is not None:
            output = self.norm(output)
        coeffs = torch.cat(coefficients, dim=0) if coefficients else \
            torch.empty((0, B, self.num_coefficients), device=src.device)
        return output, attn, coeffs.permute([1, 0, 2])


class GlobalAvg1D(nn.Module):
    """models.py:586-595 -- masked mean over the padded node axis, one kernel."""

    def forward(self, x, mask=None):
        if mask is None:
            return x.mean(dim=1)
        return ops.masked_mean(x, mask)


def _regularisation(coeff):
    """models.py:727-742 (kept as plain PyTorch; off by default, regularization=0.0)."""
    gm = torch.bmm(coeff, coeff.permute([0, 2, 1]))
    mask = 1. - torch.eye(coeff.shape[1], device=coeff.device).unsqueeze(0).repeat([coeff.shape[0], 1, 1])
    gm = gm * mask
    v1 = torch.norm(coeff, p=2, dim=[2])
    norm_mat = torch.bmm(v1.unsqueeze(-1), v1.unsqueeze(1))
    reg = torch.div(gm, norm_mat)
    reg = torch.max(torch.max(reg, dim=1).values, dim=1).values
    return torch.sum(reg)


class DiffGraphTransformerGenGCN(nn.Module):
    """models.py:487-551 (graph-level head: MUTAG / ZINC)."""

    def __init__(self, in_size, nb_class, d_model, nb_heads, dim_feedforward=2048, dropout=0.1,
                 nb_layers=4, batch_norm=False, lap_pos_enc=False, lap_pos_enc_dim=0, filter_order=4,
                 gnn_type='ChebConvDynamic', last_layer_filter=True,
                 learn_only_filter_order_coeff=False, **layer_kw):
        super().__init__()
        self.lap_pos_enc = lap_pos_enc
        self.lap_pos_enc_dim = lap_pos_enc_dim
        if lap_pos_enc and lap_pos_enc_dim > 0:
            self.embedding_lap_pos_enc = Linear(lap_pos_enc_dim, d_model)
        self.embedding = Linear(in_features=in_size, out_features=d_model, bias=False)
        encoder_layer = DiffTransformerEncoderLayer(d_model, nb_heads, dim_feedforward, dropout,
                                                    batch_norm=batch_norm, **layer_kw)
        self.encoder = DiffTransformerEncoderGenGCN(
            d_model, nb_heads, encoder_layer, nb_layers, num_coefficients=filter_order,
            gnn_type=gnn_type, last_layer_filter=last_layer_filter,
            learn_only_filter_order_coeff=learn_only_filter_order_coeff)
        self.gcn = GCNConv(d_model, d_model)                                        # :508 (unused)
        self.pooling = GlobalAvg1D()
        self.classifier = nn.Sequential(Linear(d_model, d_model), nn.ReLU(True),
                                        Linear(d_model, nb_class))

    def _embed(self, x, x_lap_pos_enc):
        output = self.embedding(x.permute(1, 0, 2))
        if self.lap_pos_enc and x_lap_pos_enc is not None:
            output = output + self.embedding_lap_pos_enc(x_lap_pos_enc.transpose(0, 1))
        return output

    def regularisation(self, coeff):
        return _regularisation(coeff)

    def forward(self, x, edge_index, batch, feature_indices, masks, pe, x_lap_pos_enc=None,
                degree=None, regularization=0.0, return_filter_coeff=False):
        output = self._embed(x, x_lap_pos_enc)
        output, attn, filter_coeff = self.encoder(output, pe, edge_index, feature_indices, batch,
                                                  degree=degree, src_key_padding_mask=masks)
        output = output.permute(1, 0, 2)
        output_pooled = self.pooling(output, masks)                                 # :532
        filter_coeff_reg = self.regularisation(filter_coeff) if regularization > 0 else 0
        if return_filter_coeff:
            return self.classifier(output_pooled), filter_coeff_reg, filter_coeff
        return self.classifier(output_pooled), filter_coeff_reg


def _graph_head_forward_static(self, x, edge_index, masks, pe, x_lap_pos_enc=None, degree=None):
    """Static-shape forward of the graph-level heads (same numbers as forward(); see
    DiffTransformerEncoderGenGCN.forward_static).  Returns the classifier output only."""
    output = self._embed(x, x_lap_pos_enc)
    output, attn, filter_coeff = self.encoder.forward_static(output, pe, edge_index, degree, masks)
    return self.classifier(self.pooling(output.permute(1, 0, 2), masks))


DiffGraphTransformerGenGCN.forward_static = _graph_head_forward_static


class DiffGraphTransformerGenGCNSBM(nn.Module):
    """models.py:1008-1076 (node-level head: PATTERN / CLUSTER)."""

    def __init__(self, in_size, nb_class, d_model, nb_heads, dim_feedforward=2048, dropout=0.1,
                 nb_layers=4, batch_norm=False, lap_pos_enc=False, lap_pos_enc_dim=0, filter_order=4,
                 gnn_type='ChebConvDynamic', last_layer_filter=True,
                 learn_only_filter_order_coeff=False, **layer_kw):
        super().__init__()
        self.lap_pos_enc = lap_pos_enc
        self.lap_pos_enc_dim = lap_pos_enc_dim
        if lap_pos_enc and lap_pos_enc_dim > 0:
            self.embedding_lap_pos_enc = Linear(lap_pos_enc_dim, d_model)
        self.embedding = Linear(in_features=in_size, out_features=d_model, bias=False)
        encoder_layer = DiffTransformerEncoderLayer(d_model, nb_heads, dim_feedforward, dropout,
                                                    batch_norm=batch_norm, **layer_kw)
        self.encoder = DiffTransformerEncoderGenGCN(
            d_model, nb_heads, encoder_layer, nb_layers, num_coefficients=filter_order,
            gnn_type=gnn_type, last_layer_filter=last_layer_filter,
            learn_only_filter_order_coeff=learn_only_filter_order_coeff)
        self.classifier = nn.Sequential(Linear(d_model, d_model), nn.ReLU(True),
                                        Linear(d_model, nb_class))

    _embed = DiffGraphTransformerGenGCN._embed

    def regularisation(self, coeff):
        return _regularisation(coeff)

    def forward(self, x, edge_index, batch, feature_indices, masks, pe, x_lap_pos_enc=None,
                degree=None, regularization=0.0, return_filter_coeff=False):
        output = self._embed(x, x_lap_pos_enc)
        output, attn, filter_coeff = self.encoder(output, pe, edge_index, feature_indices, batch,
                                                  degree=degree, src_key_padding_mask=masks)
        output = output.permute(1, 0, 2)
        filter_coeff_reg = self.regularisation(filter_coeff) if regularization > 0 else 0
        cls_output = self.classifier(output)                                        # :1069
        # cls_output[~masks] (:1070-1071) == rows listed by feature_indices, without the
        # nonzero() synchronisation boolean indexing implies
        cls_output = ops.gather_rows(cls_output, feature_indices)
        if return_filter_coeff:
            return cls_output, filter_coeff_reg, filter_coeff
        return cls_output, filter_coeff_reg


def _node_head_forward_static(self, x, edge_index, masks, pe, x_lap_pos_enc=None, degree=None):
    """Static-shape forward of the node-level head: logits for EVERY padded slot [B, Nmax, C];
    the caller's loss ignores padded slots (labels = -100) instead of gathering ``[~masks]``."""
    output = self._embed(x, x_lap_pos_enc)
    output, attn, filter_coeff = self.encoder.forward_static(output, pe, edge_index, degree, masks)
    return self.classifier(output.permute(1, 0, 2))


DiffGraphTransformerGenGCNSBM.forward_static = _node_head_forward_static


class AtomEncoder(nn.Module):
    """ogb.graphproppred.mol_encoder.AtomEncoder (un-vendored dependency; call site
    models.py:619,646): one embedding per integer atom feature, xavier-uniform, summed.
    Same state_dict keys (``atom_embedding_list.{i}.weight``)."""
    FULL_ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]

    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        for dim in self.FULL_ATOM_FEATURE_DIMS:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)

    def forward(self, x):
        out = 0
        for i in range(x.shape[1]):
            out = out + self.atom_embedding_list[i](x[:, i])
        return out


class BondEncoder(nn.Module):
    """ogb BondEncoder shell (constructed at models.py:621, never used in forward)."""
    FULL_BOND_FEATURE_DIMS = [5, 6, 2]

    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = nn.ModuleList()
        for dim in self.FULL_BOND_FEATURE_DIMS:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)

    def forward(self, edge_attr):
        out = 0
        for i in range(edge_attr.shape[1]):
            out = out + self.bond_embedding_list[i](edge_attr[:, i])
        return out


class DiffGraphTransformerGenGCNMolHiv(nn.Module):
    """models.py:598-725 (ogbg-molhiv head)."""

    def __init__(self, in_size, nb_class, d_model, nb_heads, dim_feedforward=2048, dropout=0.1,
                 nb_layers=4, batch_norm=False, lap_pos_enc=False, lap_pos_enc_dim=0, filter_order=4,
                 gnn_type='ChebConvDynamic', last_layer_filter=True,
                 learn_only_filter_order_coeff=False, use_skip_conn=True, use_default_encoder=False,
                 **layer_kw):
        super().__init__()
        if use_default_encoder:
            raise NotImplementedError("use_default_encoder=True selects the non-GenGCN encoder, "
                                      "which is outside the hot path")
        self.lap_pos_enc = lap_pos_enc
        self.lap_pos_enc_dim = lap_pos_enc_dim
        if lap_pos_enc and lap_pos_enc_dim > 0:
            self.embedding_lap_pos_enc = Linear(lap_pos_enc_dim, d_model)
        self.d_model = d_model
        self.embedding = AtomEncoder(emb_dim=d_model)
        self.edge_embeddings = BondEncoder(emb_dim=d_model)
        encoder_layer = DiffTransformerEncoderLayer(d_model, nb_heads, dim_feedforward, dropout,
                                                    batch_norm=batch_norm, **layer_kw)
        self.encoder = DiffTransformerEncoderGenGCN(
            d_model, nb_heads, encoder_layer, nb_layers, num_coefficients=filter_order,
            gnn_type=gnn_type, last_layer_filter=last_layer_filter,
            learn_only_filter_order_coeff=learn_only_filter_order_coeff, use_skip_conn=use_skip_conn)
        self.gcn = GCNConv(d_model, d_model)
        self.pooling = GlobalAvg1D()
        self.classifier = nn.Sequential(Linear(d_model, d_model), nn.LeakyReLU(True),
                                        Linear(d_model, nb_class))
        self.sigmoid = nn.Sigmoid()

    def regularisation(self, coeff):
        return _regularisation(coeff)

    def forward(self, x, edge_index, batch, feature_indices, masks, pe, x_lap_pos_enc=None,
                degree=None, regularization=0.0, return_filter_coeff=False):
        x_t = x.reshape([-1, x.shape[-1]])                                          # :645
        output = self.embedding(x_t.to(torch.int64))
        output = output.reshape([x.shape[0], x.shape[1], self.d_model]).permute(1, 0, 2)
        if self.lap_pos_enc and x_lap_pos_enc is not None:
            output = output + self.embedding_lap_pos_enc(x_lap_pos_enc.transpose(0, 1))
        output, attn, filter_coeff = self.encoder(output, pe, edge_index, feature_indices, batch,
                                                  degree=degree, src_key_padding_mask=masks)
        output = output.permute(1, 0, 2)
        output_pooled = self.pooling(output, masks)
        filter_coeff_reg = self.regularisation(filter_coeff) if regularization > 0 else 0
        cls_out = self.classifier(output_pooled)
        if return_filter_coeff:
            return cls_out.squeeze(), filter_coeff_reg, self.sigmoid(cls_out).squeeze(), filter_coeff
        return cls_out.squeeze(), filter_coeff_reg, self.sigmoid(cls_out).squeeze()


def _molhiv_forward_static(self, x, edge_index, masks, pe, x_lap_pos_enc=None, degree=None):
    x_t = x.reshape([-1, x.shape[-1]])
    output = self.embedding(x_t.to(torch.int64))
    output = output.reshape([x.shape[0], x.shape[1], self.d_model]).permute(1, 0, 2)
    if self.lap_pos_enc and x_lap_pos_enc is not None:
        output = output + self.embedding_lap_pos_enc(x_lap_pos_enc.transpose(0, 1))
    output, attn, filter_coeff = self.encoder.forward_static(output, pe, edge_index, degree, masks)
    return self.classifier(self.pooling(output.permute(1, 0, 2), masks)).squeeze(-1)


DiffGraphTransformerGenGCNMolHiv.forward_static = _molhiv_forward_static
