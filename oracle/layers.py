"""CPU restatement of the kernel-biased attention layer -- TEST INFRASTRUCTURE ONLY.

``DiffTransformerEncoderLayer`` is imported by /root/reference/transformer/models.py:4
but its source is ABSENT from the reference tree (transformer/layers.py is a
byte copy of gckn/layers.py -- SURVEY.md F1).  PARITY UNPINNED.  What is
restated here is the call-site contract (models.py:166-167, :92-93, :505-506)
plus the semantics of the upstream the reference credits (GraphiT,
README.md:129 of the reference), cross-checked against the reference's own DGL
ports (LSPE/layers/graphit_gt_layer.py:95-131,164 -- renormalise after the
kernel product with a 1e-6 guard; LSPE/layers/graphit_spectra_lspe_layer.py:505-512
-- the per-head outputs handed to the filter are the pre-``O`` projections):

    q, k, v = in_proj(src);  S = (q * dh^-1/2) k^T;  key padding -> -inf
    E = exp(S - rowmax(S)) * pe;  P = E / clamp(rowsum(E), 1e-6);  O = P v
    src2 = out_proj(concat_heads(O)) * degree;  src = norm1(src + src2)
    src = norm2(src + linear2(relu(linear1(src))))
    returns (src, P [B,H,N,N], O [B,N,H,dh])  with need_heads=True

Open choices are flags with the defaults used everywhere in this repo:
``share_qk=False`` (GraphiT's paper ties W_K to W_Q; north_star writes QK^T),
``attn_bias=False`` (GraphiT builds its attention with bias=False).
"""
import torch
from torch import nn
import torch.nn.functional as F


class OracleDiffMultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=False, share_qk=False):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.dropout = dropout
        self.share_qk = share_qk
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        if bias:
            self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        else:
            self.register_parameter('in_proj_bias', None)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        nn.init.xavier_uniform_(self.in_proj_weight)
        if bias:
            nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, src, pe=None, key_padding_mask=None, zero_padded_queries=False):
        N, B, E = src.shape
        H, dh = self.num_heads, self.head_dim
        scaling = float(dh) ** -0.5
        q, k, v = F.linear(src, self.in_proj_weight, self.in_proj_bias).chunk(3, dim=-1)
        if self.share_qk:
            k = q
        q = q * scaling
        q = q.contiguous().view(N, B * H, dh).transpose(0, 1)
        k = k.contiguous().view(N, B * H, dh).transpose(0, 1)
        v = v.contiguous().view(N, B * H, dh).transpose(0, 1)
        w = torch.bmm(q, k.transpose(1, 2))                                 # [B*H, N, N]
        if key_padding_mask is not None:
            w = w.view(B, H, N, N).masked_fill(
                key_padding_mask.unsqueeze(1).unsqueeze(2), float('-inf')).view(B * H, N, N)
        max_val = w.max(dim=-1, keepdim=True)[0]
        w = torch.exp(w - max_val)
        w = w.view(B, H, N, N)
        if pe is not None:
            w = w * pe.unsqueeze(1)
        w = w / w.sum(dim=-1, keepdim=True).clamp(min=1e-6)
        if zero_padded_queries and key_padding_mask is not None:
            w = w.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(3), 0.0)
        w = F.dropout(w, p=self.dropout, training=self.training)
        o = torch.bmm(w.view(B * H, N, N), v)                               # [B*H, N, dh]
        heads = o.view(B, H, N, dh).permute(0, 2, 1, 3)                     # [B, N, H, dh]
        o = o.transpose(0, 1).contiguous().view(N, B, E)
        o = self.out_proj(o)
        return o, w, heads


class OracleDiffTransformerEncoderLayer(nn.Module):
    """Constructor contract of models.py:505-506:
    ``(d_model, nb_heads, dim_feedforward, dropout, batch_norm=...)``."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu",
                 batch_norm=False, attn_bias=False, share_qk=False):
        super().__init__()
        self.self_attn = OracleDiffMultiheadAttention(d_model, nhead, dropout=dropout,
                                                      bias=attn_bias, share_qk=share_qk)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.batch_norm = batch_norm
        if batch_norm:
            self.norm1 = nn.BatchNorm1d(d_model)
            self.norm2 = nn.BatchNorm1d(d_model)
        else:
            self.norm1 = nn.LayerNorm(d_model)
            self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        assert activation == "relu"
        self.scaling = None
        self.zero_padded_queries = False

    def forward(self, src, pe=None, degree=None, src_mask=None, src_key_padding_mask=None,
                need_heads=False):
        assert src_mask is None, "attn_mask is never passed by the reference (models.py:166)"
        src2, attn, heads = self.self_attn(src, pe=pe, key_padding_mask=src_key_padding_mask,
                                           zero_padded_queries=self.zero_padded_queries)
        if degree is not None:
            src2 = degree.transpose(0, 1).contiguous().unsqueeze(-1) * src2
        else:
            if self.scaling is None:
                self.scaling = 1. / pe.diagonal(dim1=1, dim2=2).max().item()
            src2 = (self.scaling * pe.diagonal(dim1=1, dim2=2)).transpose(0, 1) \
                .contiguous().unsqueeze(-1) * src2
        src = src + self.dropout1(src2)
        if self.batch_norm:
            bsz = src.shape[1]
            src = src.reshape(-1, src.shape[-1])
        src = self.norm1(src)
        src2 = self.linear2(self.dropout(F.relu(self.linear1(src))))
        src = src + self.dropout2(src2)
        src = self.norm2(src)
        if self.batch_norm:
            src = src.view(-1, bsz, src.shape[-1])
        if need_heads:
            return src, attn, heads
        return src, attn
