"""Drop-in ``DiffTransformerEncoderLayer`` -- the kernel-biased attention layer.

The reference imports this class from ``transformer/layers.py`` (transformer/models.py:4) but its
source is absent from the tree, so the contract is the call site
(models.py:166-167, :92-93, :505-506) plus the upstream GraphiT semantics the reference credits
(README.md:129).  Parameter names follow ``nn.TransformerEncoderLayer`` so reference checkpoints
load: ``self_attn.in_proj_weight``, ``self_attn.out_proj.weight``, ``linear1/2``, ``norm1/2``.

The attention core -- S = scale*QK^T, key-padding mask, exp(S - rowmax) * pe, renormalise with a
1e-6 clamp, O = P V, per head -- is ONE fused sm_100a kernel (csrc/attention.cu) that also emits
the per-head attention matrix and per-head outputs the FeTA encoder consumes.  Projections, FFN
and norms are plain library GEMMs / PyTorch.
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import ops


class Linear(nn.Linear):
    """``nn.Linear`` (same parameters / state_dict keys) whose weight and bias gradients are reduced
    over the token axis by csrc/dense.cu instead of a 54-CTA SIMT sgemm."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class DiffMultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=False, share_qk=False):
        super().__init__()
        if embed_dim % num_heads != 0:
            raise ValueError("embed_dim must be divisible by num_heads")
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.dropout = dropout
        self.share_qk = share_qk
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        if bias:
            self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        else:
            self.register_parameter('in_proj_bias', None)
        self.out_proj = Linear(embed_dim, embed_dim, bias=bias)
        nn.init.xavier_uniform_(self.in_proj_weight)
        if bias:
            nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, src, pe=None, key_padding_mask=None, with_residual=False, drop=None, defer_out_proj=False,
                need_attn=True):
        """src [Nmax, B, d] -> (out [Nmax, B, d], attn [B, H, Nmax, Nmax], heads [B, Nmax, H, dh]).
        ``with_residual``: a 4th output, ``src`` routed through the in-projection's autograd node (its
        gradient is then folded into the in-projection's dX GEMM, ops.LinearFn)."""
        N, B, E = src.shape
        if drop is None and self.dropout > 0.0 and self.training:
            # attention-weight dropout (F.dropout on P before P V): the multipliers come from torch's Philox stream
            # (CUDA-graph safe), the kernels apply them in the forward pass and again in the backward pass
            drop = ops.dropout_multiplier((B, self.num_heads, N, N), self.dropout, src.device)
        res = None
        if with_residual:
            qkv, res = ops.linear_res(src, self.in_proj_weight, self.in_proj_bias)
        else:
            qkv = ops.linear(src, self.in_proj_weight, self.in_proj_bias)    # library GEMM (+ own wgrad)
        attn, o_sf = ops.diff_attention(qkv, pe, key_padding_mask, self.num_heads,
                                        float(self.head_dim) ** -0.5, self.share_qk, drop=drop,
                                        need_attn=need_attn)
        # the kernel writes O seq-first, so concat-heads -> out_proj needs no copy; `heads` is the
        # [B, Nmax, H, dh] view the FeTA encoder consumes (models.py:179)
        # defer_out_proj: the caller fuses out_proj with the degree scale, the residual and norm1 (one launch)
        out = o_sf.view(N, B, E) if defer_out_proj else self.out_proj(o_sf.view(N, B, E))
        if with_residual:
            return out, attn, o_sf.permute(1, 0, 2, 3), res
        return out, attn, o_sf.permute(1, 0, 2, 3)


class DiffTransformerEncoderLayer(nn.Module):
    """ctor contract models.py:505-506: ``(d_model, nb_heads, dim_feedforward, dropout, batch_norm=)``;
    call contract models.py:166-167 / :92-93."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu",
                 batch_norm=False, attn_bias=False, share_qk=False):
        super().__init__()
        self.self_attn = DiffMultiheadAttention(d_model, nhead, dropout=dropout, bias=attn_bias,
                                                share_qk=share_qk)
        self.linear1 = Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = Linear(dim_feedforward, d_model)
        self.batch_norm = batch_norm
        if batch_norm:
            self.norm1 = nn.BatchNorm1d(d_model)
            self.norm2 = nn.BatchNorm1d(d_model)
        else:
            self.norm1 = nn.LayerNorm(d_model)
            self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        if activation != "relu":
            raise NotImplementedError("only relu")
        self.scaling = None

    def forward(self, src, pe=None, degree=None, src_mask=None, src_key_padding_mask=None,
                need_heads=False, rowscale=None, bn_rows=None, need_attn=True, after_attention=None):
        """``rowscale`` (extension): the seq-first ``degree.t().contiguous()`` precomputed once per forward by
        the encoder instead of once per layer.  ``bn_rows`` (extension, BatchNorm variant): 0/1 weight per
        flattened row ``[Nmax * B]``; rows with 0 stay out of the batch statistics (static-shape batches).
        ``need_attn=False`` (extension): the caller does not read the attention matrix of this layer; it comes back as
        None and is never materialised (matrix-free attention kernels).
        ``after_attention`` (extension): ``callback(attn, out_each_head)`` invoked as soon as the attention core has
        been issued, before the layer's token-wise tail -- the encoder starts the filter branch there (on another
        stream), which depends on the attention outputs only."""
        if src_mask is not None:
            raise NotImplementedError("src_mask (attn_mask) is never passed by the reference "
                                      "(models.py:166) and is not implemented")
        fused = not self.batch_norm
        dm = src.shape[-1]
        no_drop = (not self.training) or self.dropout1.p == 0.0
        fuse_ln = fused and no_drop and src.is_cuda and ops.linear_layernorm_enabled(dm, dm) \
            and ops.linear_layernorm_enabled(self.linear1.weight.shape[0], dm)
        if fused:
            src2, attn, heads, src = self.self_attn(src, pe=pe, key_padding_mask=src_key_padding_mask,
                                                    with_residual=True, defer_out_proj=fuse_ln, need_attn=need_attn)
        else:
            src2, attn, heads = self.self_attn(src, pe=pe, key_padding_mask=src_key_padding_mask,
                                               need_attn=need_attn)
        if after_attention is not None:
            after_attention(attn, heads)
        if rowscale is not None:
            pass
        elif degree is not None:
            rowscale = degree.transpose(0, 1).contiguous()                      # [Nmax, B]
        else:
            if pe is None:
                raise ValueError("DiffTransformerEncoderLayer needs `degree` or `pe`")
            if self.scaling is None:
                self.scaling = 1. / pe.diagonal(dim1=1, dim2=2).max().item()
            rowscale = (self.scaling * pe.diagonal(dim1=1, dim2=2)).transpose(0, 1).contiguous()
        if self.batch_norm and self.training and ops.batchnorm_supported(src.shape[-1]) and src.is_cuda \
                and self.dropout1.p == 0.0:
            # fused residual + BatchNorm1d over the flattened rows (padding rows included, like the reference)
            bsz, dm = src.shape[1], src.shape[-1]
            src = ops.add_batch_norm(src.reshape(-1, dm), src2.reshape(-1, dm), self.norm1,
                                     bscale=rowscale.reshape(-1), roww=bn_rows)
            h = ops.linear(src, self.linear1.weight, self.linear1.bias, relu=True)
            src2 = ops.linear(self.dropout(h), self.linear2.weight, self.linear2.bias)
            src = ops.add_batch_norm(src, src2, self.norm2, roww=bn_rows)
            src = src.view(-1, bsz, dm)
        elif self.batch_norm:
            if bn_rows is not None:
                raise NotImplementedError("bn_rows needs the fused BatchNorm path (training mode, d_model dividing 256)")
            src = src + self.dropout1(rowscale.unsqueeze(-1) * src2)
            bsz = src.shape[1]
            src = src.reshape(-1, src.shape[-1])
            src = self.norm1(src)
            src2 = self.linear2(self.dropout(F.relu(self.linear1(src))))
            src = src + self.dropout2(src2)
            src = self.norm2(src)
            src = src.view(-1, bsz, src.shape[-1])
        elif fuse_ln:
            # tcgen05 path, 4 launches per layer forward: in-projection, attention, [out_proj + degree scale +
            # residual + norm1], linear1(+ReLU), [linear2 + residual + norm2]
            op = self.self_attn.out_proj
            src = ops.linear_add_layer_norm(src2, op.weight, op.bias, src, self.norm1.weight, self.norm1.bias,
                                            self.norm1.eps, bscale=rowscale.reshape(-1))
            h, src = ops.linear_res(src, self.linear1.weight, self.linear1.bias, relu=True, grad_premasked=True)
            src = ops.linear_add_layer_norm(h, self.linear2.weight, self.linear2.bias, src, self.norm2.weight,
                                            self.norm2.bias, self.norm2.eps, mask_input_grad=True)
        else:                                    # residual add fused into the LayerNorm kernels
            src = ops.add_layer_norm(src, self.dropout1(src2), self.norm1.weight, self.norm1.bias, self.norm1.eps,
                                     bscale=rowscale.reshape(-1))       # degree * src2 fused in
            # ReLU in linear1's epilogue; with the tensor-core GEMMs its backward mask moves into linear2's dX
            # epilogue (otherwise linear1's backward applies it, one threshold kernel)
            dff, dm = self.linear1.weight.shape
            fm = ops.linear_tc_enabled(dm, dff) and ops.linear_tc_enabled(dff, dm)
            h, src = ops.linear_res(src, self.linear1.weight, self.linear1.bias, relu=True, grad_premasked=fm)
            src2 = ops.linear(self.dropout(h), self.linear2.weight, self.linear2.bias, mask_input_grad=fm)
            src = ops.add_layer_norm(src, self.dropout2(src2), self.norm2.weight, self.norm2.bias, self.norm2.eps)
        if need_heads:
            return src, attn, heads
        return src, attn
