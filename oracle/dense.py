"""Independent dense-matrix cross-check of the oracle -- TEST INFRASTRUCTURE ONLY.

Written from the *mathematical* definition (docstring of ChebNetDynamic.py:30-78
and SURVEY.md section 8 A1/A2/A4), not from the reference's op sequence, so the
sparse restatements in ``oracle/cheb.py`` / ``oracle/encoder.py`` are not
checked against themselves.
"""
import torch


def dense_scaled_laplacian(edge_index, num_nodes, dtype=torch.float64):
    """L_hat = 2 L_sym / lambda_max - I with lambda_max = 2  ==  -D^-1/2 A D^-1/2.

    A[t, s] counts edges s->t after dropping self loops (multi-edges add up);
    D is the *out*-degree (edges counted at their source, as PyG get_laplacian
    does); isolated nodes get D^-1/2 := 0.
    """
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    src, dst = src[keep], dst[keep]
    A = torch.zeros(num_nodes, num_nodes, dtype=dtype)
    A.index_put_((dst, src), torch.ones(src.numel(), dtype=dtype), accumulate=True)
    deg = torch.zeros(num_nodes, dtype=dtype).index_add_(0, src, torch.ones(src.numel(), dtype=dtype))
    dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
    return -(dis.view(-1, 1) * A * dis.view(1, -1))


def dense_cheb(x, edge_index, theta, graph_of_row, bias=None):
    """out[i] = sum_k T_k[i] @ theta[k, g(i)] (+ bias), T_k by dense products."""
    R = x.size(0)
    L = dense_scaled_laplacian(edge_index, R, dtype=x.dtype)
    K = theta.size(0)
    Ts = [x]
    if K > 1:
        Ts.append(L @ x)
    for _ in range(2, K):
        Ts.append(2.0 * (L @ Ts[-1]) - Ts[-2])
    th = theta[:, graph_of_row.long()]                      # [K, R, Fin, Fout]
    out = sum(torch.einsum('ri,rio->ro', Ts[k], th[k]) for k in range(K))
    if bias is not None:
        out = out + bias
    return out


def dense_coeff_scalar(attn_g):
    """Closed form of the all-ones GCN of models.py:280-282 for ONE (head, graph).

    attn_g [n, n] (row i = query/source, column j = key/target).  Returns s [n]
    with GCNConv(ones)[j] = s[j] * colsum(W) + b.
    """
    n = attn_g.size(0)
    a = attn_g.clone()
    diag = torch.diagonal(a).clone()
    loop = torch.where(diag != 0, diag, torch.ones_like(diag))
    a.fill_diagonal_(0)
    a = a + torch.diag(loop)
    deg = a.sum(dim=0)                                      # over sources i, per target j
    dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
    return (dis.view(-1, 1) * a * dis.view(1, -1)).sum(dim=0)
