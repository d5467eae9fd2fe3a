"""Generates the frozen golden vectors under tests/golden/ FROM THE ORACLE (fixed seeds).

The reference ships no golden vectors and cannot be imported here, so these
pin the oracle against accidental drift and give the GPU tests a reference that does not need
the oracle code path at all.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import make_batch, random_batch_graph  # noqa: E402
from oracle.cheb import OracleChebConvDynamic  # noqa: E402
from oracle.layers import OracleDiffTransformerEncoderLayer  # noqa: E402
import oracle.models as omodels  # noqa: E402
from feta_tmlr_b200 import synthetic  # noqa: E402


def cheb_case():
    sizes = [5, 7, 1, 4, 12]
    ei, batch, R = random_batch_graph(11, sizes, directed_extra=2)
    F, K, G = 8, 4, len(sizes)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(R, F, generator=g).requires_grad_()
    coeff = (torch.randn(G, K * F * F, generator=g) * 0.3).requires_grad_()
    bias = torch.randn(F, generator=g)
    dout = torch.randn(R, F, generator=g)
    m = OracleChebConvDynamic(F, F, K)
    m.bias.data.copy_(bias)
    out = m(x, ei, coeff.reshape(-1, K, F, F).permute(1, 0, 2, 3), batch=batch.float())
    out.backward(dout)
    return dict(edge_index=ei, batch=batch, x=x.detach(), coeff=coeff.detach(), bias=bias, dout=dout,
                out=out.detach(), dx=x.grad, dcoeff=coeff.grad, dbias=m.bias.grad, F=F, K=K)


def attention_case():
    torch.manual_seed(12)
    d, H, B, nmax = 32, 4, 3, 9
    layer = OracleDiffTransformerEncoderLayer(d, H, 2 * d, 0.0)
    layer.zero_padded_queries = True
    g = torch.Generator().manual_seed(12)
    lens = torch.tensor([9, 4, 6])
    mask = torch.arange(nmax)[None, :] >= lens[:, None]
    src = torch.randn(nmax, B, d, generator=g).requires_grad_()
    a = torch.rand(B, nmax, nmax, generator=g)
    pe = (a + a.transpose(1, 2)) * 0.5 * ((~mask)[:, :, None] & (~mask)[:, None, :])
    degree = torch.rand(B, nmax, generator=g) * (~mask)
    out, attn, heads = layer(src, pe=pe, degree=degree, src_key_padding_mask=mask, need_heads=True)
    w = torch.randn(out.shape, generator=g)
    wh = torch.randn(heads.shape, generator=g)
    ((out * w).sum() + (heads * wh).sum()).backward()
    return dict(state_dict=layer.state_dict(), src=src.detach(), pe=pe, degree=degree, mask=mask, out=out.detach(),
                attn=attn.detach(), heads=heads.detach(), w=w, wh=wh, dsrc=src.grad, d=d, H=H)


def model_case():
    cfg, graphs, store, batch = make_batch("MUTAG", 4, seed=13)
    torch.manual_seed(13)
    m = synthetic.build_model("MUTAG", omodels, layers=2, d_model=8, heads=2)
    for layer in m.encoder.layers:
        layer.zero_padded_queries = True
    px, mask, pe, lap, deg, labels, ei, bi, fi = batch[:9]
    out, _, coeff = m(px, ei, bi, fi, mask, pe, lap, deg, return_filter_coeff=True)   # literal all-pairs GCN
    loss = torch.nn.functional.cross_entropy(out, labels.long())
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    return dict(state_dict=m.state_dict(), batch=[t for t in batch[:9]], out=out.detach(), coeff=coeff.detach(),
                loss=loss.detach(), grads=grads, over=dict(layers=2, d_model=8, heads=2))


if __name__ == "__main__":
    torch.save(cheb_case(), os.path.join(HERE, "cheb_case.pt"))
    torch.save(attention_case(), os.path.join(HERE, "attention_case.pt"))
    torch.save(model_case(), os.path.join(HERE, "model_case.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
