"""CPU restatement of the FeTA encoder + model heads -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/transformer/models.py:103-368 (DiffTransformerEncoderGenGCN),
:487-551 (DiffGraphTransformerGenGCN), :586-595 (GlobalAvg1D), :598-725 (MolHiv
head) and :1008-1076 (SBM head).  PARITY UNPINNED (no reference tests).
``utils.DEVICE`` of the reference is replaced by the input tensor's device.
"""
import copy
import math

import numpy as np
import torch
from torch import nn

from . import pyg17
from .cheb import OracleChebConvDynamic
from .layers import OracleDiffTransformerEncoderLayer


class OracleGCNConv(nn.Module):
    """PyG-1.7 GCNConv parameter shell: ``weight [in, out]`` (glorot), ``bias`` (zeros)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        stdv = math.sqrt(6.0 / (in_channels + out_channels))
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, edge_index, edge_weight=None):
        return pyg17.gcn_conv(x, edge_index, edge_weight, self.weight, self.bias)


class OracleEncoderGenGCN(nn.Module):
    """``DiffTransformerEncoderGenGCN`` -- models.py:103-368 (gnn_type 'ChebConvDynamic' / 'ARMAConvDynamic')."""

    def __init__(self, d_model, num_heads, encoder_layer, num_layers, norm=None,
                 num_coefficients=4, laplacian_norm='sym', gnn_type='ChebConvDynamic',
                 last_layer_filter=True, learn_only_filter_order_coeff=False,
                 use_skip_conn=True):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        assert gnn_type in ('ChebConvDynamic', 'ARMAConvDynamic')
        self.num_coefficients = num_coefficients
        self.order = num_coefficients                                                  # :127,130
        dh = d_model // num_heads
        if gnn_type == 'ARMAConvDynamic':                                              # :135-139
            from .arma import OracleARMAConvDynamic
            self.num_coefficients = self.num_coefficients * 2
            self.spectral_gnns = OracleARMAConvDynamic(dh, dh, num_stacks=self.order, num_layers=1)
        elif learn_only_filter_order_coeff:
            self.spectral_gnns = OracleChebConvDynamic(dh, dh, self.num_coefficients,
                                                       normalization=laplacian_norm,
                                                       learn_only_filter_order_coeff=True)
        else:
            self.filter_in_channels = dh
            self.filter_out_channels = dh
            self.num_coefficients = self.order * dh * dh                               # :133
            self.spectral_gnns = OracleChebConvDynamic(dh, dh, self.order,
                                                       normalization=laplacian_norm)
        self.gcn = OracleGCNConv(self.num_coefficients, self.num_coefficients)          # :144
        self.linear = nn.Linear(self.num_coefficients, self.num_coefficients)          # :145
        self.linear_cat = nn.Linear(2 * d_model, d_model)                              # :146
        self.gnn_type = gnn_type
        self.num_heads = num_heads
        self.last_layer_filter = last_layer_filter
        self.learn_only_filter_order_coeff = learn_only_filter_order_coeff
        self.use_skip_conn = use_skip_conn
        self.collapsed_coeff = False   # True: closed form of :280-283 (same numbers, no 15 GB tensor)

    # models.py:240-287
    def get_filter_coefficients(self, attn_weights, edge_index, feature_indices, batch, masks):
        dev = attn_weights.device
        B, H = attn_weights.shape[0], attn_weights.shape[1]
        masks = masks.repeat([H, 1])                                                   # :244
        inv_masks = ~masks
        g_len_list = torch.sum(inv_masks, dim=-1).detach().cpu().numpy()               # :246
        if self.collapsed_coeff:
            return self._coeff_collapsed(attn_weights, g_len_list)
        masked_indices, edge_index, batch = [], [], []
        node_offset = 0
        for b_idx in range(len(inv_masks)):                                            # :252-258
            g_len = int(g_len_list[b_idx])
            edge_idx = np.mgrid[node_offset:node_offset + g_len, node_offset:node_offset + g_len]
            edge_index.append(edge_idx.reshape([2, -1]))
            node_offset += g_len
            batch.append(b_idx * np.ones((g_len)))
            masked_indices.append([b_idx * np.ones((g_len)), np.arange(g_len)])
        edge_index = np.concatenate(edge_index, axis=1)
        batch = np.concatenate(batch, axis=0)
        masked_indices = np.concatenate(masked_indices, axis=1).T
        edge_index = torch.from_numpy(edge_index).to(dev)
        batch = torch.from_numpy(batch.astype(np.int64)).to(dev)
        inv_masks_int = inv_masks.to(int)                                              # :267-270
        t1 = inv_masks_int.unsqueeze(1).repeat([1, inv_masks.shape[1], 1])
        t2 = inv_masks_int.unsqueeze(1).repeat([1, inv_masks.shape[1], 1]).permute([0, 2, 1])
        t3 = t1 * t2
        edge_weight = attn_weights.permute([1, 0, 2, 3]).reshape(                      # :275
            [self.num_heads * B, attn_weights.shape[2], attn_weights.shape[3]])[t3 == 1]
        non_zero_indices = torch.where(edge_weight != 0.0)[0].detach()                 # :276
        mi = torch.from_numpy(masked_indices.astype(np.int64))
        x_c = torch.ones((B * self.num_heads, attn_weights.shape[2], self.num_coefficients),
                         dtype=attn_weights.dtype)[mi[:, 0], mi[:, 1], :].to(dev)      # :280
        edge_weight = edge_weight[non_zero_indices]                                    # :281
        x_c = torch.tanh(self.gcn(x_c, edge_index[:, non_zero_indices],                # :282
                                  edge_weight=edge_weight.detach()))
        gcn_pool = pyg17.global_mean_pool(x_c, batch)                                  # :283
        pooled_coeff = self.linear(gcn_pool)                                           # :284
        return pooled_coeff.reshape((self.num_heads, B, pooled_coeff.shape[-1]))       # :285

    def _coeff_collapsed(self, attn_weights, g_len_list):
        """Closed form of :280-283: with x == 1 the GCN output of node j is
        ``s_j * colsum(W) + b``; see oracle/dense.py::dense_coeff_scalar."""
        from .dense import dense_coeff_scalar
        B, H = attn_weights.shape[0], attn_weights.shape[1]
        wbar = self.gcn.weight.sum(dim=0)
        pooled = []
        a = attn_weights.detach().permute([1, 0, 2, 3]).reshape(H * B, *attn_weights.shape[2:])
        for g in range(H * B):
            n = int(g_len_list[g])
            s = dense_coeff_scalar(a[g, :n, :n])
            pooled.append(torch.tanh(s.view(-1, 1) * wbar.view(1, -1) + self.gcn.bias).mean(dim=0))
        pooled_coeff = self.linear(torch.stack(pooled, dim=0))
        return pooled_coeff.reshape((H, B, pooled_coeff.shape[-1]))

    # models.py:346-368
    def filter(self, filter_coeff, graph_signal, edge_index, feature_indices, batch, spectral_gnn):
        x = graph_signal[feature_indices[:, 0].long(), feature_indices[:, 1].long(), :]  # :347
        if self.gnn_type == 'ARMAConvDynamic':                                           # :361-363
            filter_coeff = filter_coeff.reshape((-1, self.order * 2))
            return spectral_gnn(x, edge_index, filter_coeff, batch=batch)
        if not self.learn_only_filter_order_coeff:                                       # :356-357
            filter_coeff = filter_coeff.reshape(
                (-1, self.order, self.filter_in_channels, self.filter_out_channels)
            ).permute([1, 0, 2, 3])
        else:
            filter_coeff = filter_coeff.reshape((-1, self.order)).permute([1, 0])        # :359
        return spectral_gnn(x, edge_index, filter_coeff, batch=batch)                    # :360

    # models.py:155-238
    def forward(self, src, pe, edge_index, feature_indices, batch, degree=None, mask=None,
                src_key_padding_mask=None):
        dev = src.device
        output = src
        num_batches = int(torch.max(batch)) + 1
        coefficients = torch.empty((0, num_batches, self.num_coefficients), device=dev)
        allout_filtered = None
        num_layers = len(self.layers)
        attn = None
        for layer_num, mod in enumerate(self.layers):
            output, attn, out_each_head = mod(output, pe=pe, degree=degree, src_mask=mask,
                                              src_key_padding_mask=src_key_padding_mask,
                                              need_heads=True)
            if self.last_layer_filter and layer_num + 1 != num_layers:                  # :169-171
                continue
            coeff_all_heads = self.get_filter_coefficients(attn, edge_index, feature_indices,
                                                           batch, src_key_padding_mask)
            coeff = coeff_all_heads.reshape((coeff_all_heads.shape[0] * coeff_all_heads.shape[1],
                                             coeff_all_heads.shape[2]))                # :178
            out_heads = out_each_head.permute([2, 0, 1, 3]).reshape(                     # :179
                (out_each_head.shape[2] * out_each_head.shape[0], out_each_head.shape[1],
                 out_each_head.shape[3]))
            batch_size = torch.max(batch) + 1
            H = self.num_heads
            batch_offset = torch.cat([torch.ones(batch.shape, device=dev) * i * batch_size
                                      for i in range(H)], dim=0)                        # :181
            batch_all_heads = batch.repeat((H)) + batch_offset                          # :182 (float)
            fi_offset = torch.cat([torch.ones((feature_indices.shape[0]), device=dev) * i * batch_size
                                   for i in range(H)], dim=0)                           # :183
            feature_indices_all_heads = feature_indices.repeat((H, 1))                  # :184
            feature_indices_all_heads[:, 0] = feature_indices_all_heads[:, 0] + fi_offset  # :185
            filtered = self.filter(coeff, out_heads, edge_index, feature_indices_all_heads,
                                   batch_all_heads, self.spectral_gnns)                 # :186 (F4:
            #   edge_index is NOT tiled per head -> only head 0's rows have edges)
            coefficients = torch.cat((coefficients, coeff_all_heads), dim=0)            # :198
            filtered_output = filtered.reshape((H, filtered.shape[0] // H, filtered.shape[1])) \
                .permute([1, 0, 2]).reshape([filtered.shape[0] // H, filtered.shape[1] * H])  # :200
            out_filtered = torch.zeros(output.shape, device=dev, dtype=output.dtype)    # :201
            out_filtered[feature_indices[:, 1], feature_indices[:, 0], :] = filtered_output  # :202
            if self.use_skip_conn:                                                      # :209-216
                allout_filtered = out_filtered if allout_filtered is None \
                    else allout_filtered + out_filtered
            else:
                allout_filtered = out_filtered
                output = allout_filtered
        if self.use_skip_conn:                                                          # :221-233
            if allout_filtered is not None:
                output = self.linear_cat(torch.cat((output, allout_filtered), dim=-1))
        else:
            if allout_filtered is not None:
                output = allout_filtered
        if self.norm is not None:
            output = self.norm(output)
        return output, attn, coefficients.permute([1, 0, 2])                            # :238


class OracleGlobalAvg1D(nn.Module):
    """models.py:586-595."""

    def forward(self, x, mask=None):
        if mask is None:
            return x.mean(dim=1)
        mask = (~mask).float().unsqueeze(-1).to(x.dtype)
        x = x * mask
        return x.sum(dim=1) / mask.sum(dim=1)


class OracleDiffGraphTransformerGenGCN(nn.Module):
    """models.py:487-551 (graph-level head: MUTAG / ZINC)."""

    def __init__(self, in_size, nb_class, d_model, nb_heads, dim_feedforward=2048, dropout=0.1,
                 nb_layers=4, batch_norm=False, lap_pos_enc=False, lap_pos_enc_dim=0,
                 filter_order=4, gnn_type='ChebConvDynamic', last_layer_filter=True,
                 learn_only_filter_order_coeff=False, **layer_kw):
        super().__init__()
        self.lap_pos_enc = lap_pos_enc
        self.lap_pos_enc_dim = lap_pos_enc_dim
        if lap_pos_enc and lap_pos_enc_dim > 0:
            self.embedding_lap_pos_enc = nn.Linear(lap_pos_enc_dim, d_model)
        self.embedding = nn.Linear(in_features=in_size, out_features=d_model, bias=False)
        encoder_layer = OracleDiffTransformerEncoderLayer(d_model, nb_heads, dim_feedforward,
                                                          dropout, batch_norm=batch_norm,
                                                          **layer_kw)
        self.encoder = OracleEncoderGenGCN(
            d_model, nb_heads, encoder_layer, nb_layers, num_coefficients=filter_order,
            gnn_type=gnn_type, last_layer_filter=last_layer_filter,
            learn_only_filter_order_coeff=learn_only_filter_order_coeff)
        self.gcn = OracleGCNConv(d_model, d_model)          # :508 (unused in forward)
        self.pooling = OracleGlobalAvg1D()
        self.classifier = nn.Sequential(nn.Linear(d_model, d_model), nn.ReLU(True),
                                        nn.Linear(d_model, nb_class))

    def embed(self, x, x_lap_pos_enc):
        output = self.embedding(x.permute(1, 0, 2))
        if self.lap_pos_enc and x_lap_pos_enc is not None:
            output = output + self.embedding_lap_pos_enc(x_lap_pos_enc.transpose(0, 1))
        return output

    def forward(self, x, edge_index, batch, feature_indices, masks, pe, x_lap_pos_enc=None,
                degree=None, regularization=0.0, return_filter_coeff=False):
        output = self.embed(x, x_lap_pos_enc)
        output, attn, filter_coeff = self.encoder(output, pe, edge_index, feature_indices, batch,
                                                  degree=degree, src_key_padding_mask=masks)
        output = output.permute(1, 0, 2)
        output_pooled = self.pooling(output, masks)
        if return_filter_coeff:
            return self.classifier(output_pooled), 0, filter_coeff
        return self.classifier(output_pooled), 0


class OracleDiffGraphTransformerGenGCNSBM(OracleDiffGraphTransformerGenGCN):
    """models.py:1008-1076 (node-level head: PATTERN / CLUSTER)."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        del self.gcn                                          # :1029 commented out in the SBM head
        del self.pooling

    def forward(self, x, edge_index, batch, feature_indices, masks, pe, x_lap_pos_enc=None,
                degree=None, regularization=0.0, return_filter_coeff=False):
        output = self.embed(x, x_lap_pos_enc)
        output, attn, filter_coeff = self.encoder(output, pe, edge_index, feature_indices, batch,
                                                  degree=degree, src_key_padding_mask=masks)
        output = output.permute(1, 0, 2)
        cls_output = self.classifier(output)                                           # :1069
        cls_output = cls_output[~masks]                                                # :1070-1071
        if return_filter_coeff:
            return cls_output, 0, filter_coeff
        return cls_output, 0


class OracleAtomEncoder(nn.Module):
    """ogb ``AtomEncoder`` (third party, un-vendored; call site models.py:619,646):
    one ``nn.Embedding`` per integer atom feature, xavier-uniform, summed."""
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


class OracleBondEncoder(nn.Module):
    """ogb ``BondEncoder`` parameter shell (constructed at models.py:621, never called)."""
    FULL_BOND_FEATURE_DIMS = [5, 6, 2]

    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = nn.ModuleList()
        for dim in self.FULL_BOND_FEATURE_DIMS:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)


class OracleDiffGraphTransformerGenGCNMolHiv(nn.Module):
    """models.py:598-725 (molhiv head; BondEncoder of :621 is constructed but unused)."""

    def __init__(self, in_size, nb_class, d_model, nb_heads, dim_feedforward=2048, dropout=0.1,
                 nb_layers=4, batch_norm=False, lap_pos_enc=False, lap_pos_enc_dim=0,
                 filter_order=4, gnn_type='ChebConvDynamic', last_layer_filter=True,
                 learn_only_filter_order_coeff=False, use_skip_conn=True, **layer_kw):
        super().__init__()
        self.lap_pos_enc = lap_pos_enc
        self.lap_pos_enc_dim = lap_pos_enc_dim
        if lap_pos_enc and lap_pos_enc_dim > 0:
            self.embedding_lap_pos_enc = nn.Linear(lap_pos_enc_dim, d_model)
        self.d_model = d_model
        self.embedding = OracleAtomEncoder(emb_dim=d_model)
        self.edge_embeddings = OracleBondEncoder(emb_dim=d_model)                      # :621 (unused)
        encoder_layer = OracleDiffTransformerEncoderLayer(d_model, nb_heads, dim_feedforward,
                                                          dropout, batch_norm=batch_norm,
                                                          **layer_kw)
        self.encoder = OracleEncoderGenGCN(
            d_model, nb_heads, encoder_layer, nb_layers, num_coefficients=filter_order,
            gnn_type=gnn_type, last_layer_filter=last_layer_filter,
            learn_only_filter_order_coeff=learn_only_filter_order_coeff,
            use_skip_conn=use_skip_conn)
        self.gcn = OracleGCNConv(d_model, d_model)
        self.pooling = OracleGlobalAvg1D()
        self.classifier = nn.Sequential(nn.Linear(d_model, d_model), nn.LeakyReLU(True),
                                        nn.Linear(d_model, nb_class))
        self.sigmoid = nn.Sigmoid()

    def forward(self, x, edge_index, batch, feature_indices, masks, pe, x_lap_pos_enc=None,
                degree=None, regularization=0.0, return_filter_coeff=False):
        x_t = x.reshape([-1, x.shape[-1]])                                             # :645
        output = self.embedding(x_t.to(int))
        output = output.reshape([x.shape[0], x.shape[1], self.d_model]).permute(1, 0, 2)
        if self.lap_pos_enc and x_lap_pos_enc is not None:
            output = output + self.embedding_lap_pos_enc(x_lap_pos_enc.transpose(0, 1))
        output, attn, filter_coeff = self.encoder(output, pe, edge_index, feature_indices, batch,
                                                  degree=degree, src_key_padding_mask=masks)
        output = output.permute(1, 0, 2)
        output_pooled = self.pooling(output, masks)
        cls_out = self.classifier(output_pooled)
        if return_filter_coeff:
            return cls_out.squeeze(), 0, self.sigmoid(cls_out).squeeze(), filter_coeff
        return cls_out.squeeze(), 0, self.sigmoid(cls_out).squeeze()
