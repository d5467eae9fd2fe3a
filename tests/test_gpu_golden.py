"""GPU parity against the committed golden vectors (no oracle code on this path)."""
import os

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


def _load(name, dev):
    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    return g


def test_golden_cheb(cuda):
    from feta_tmlr_b200 import ChebConvDynamic
    g = _load("cheb_case", cuda)
    F, K = g['F'], g['K']
    m = ChebConvDynamic(F, F, K).to(cuda)
    m.bias.data.copy_(g['bias'])
    x, c = g['x'].to(cuda).requires_grad_(), g['coeff'].to(cuda).requires_grad_()
    out = m(x, g['edge_index'].to(cuda), c.reshape(-1, K, F, F).permute(1, 0, 2, 3), batch=g['batch'].float().to(cuda))
    out.backward(g['dout'].to(cuda))
    assert rel_err(out, g['out']) < TOL and rel_err(x.grad, g['dx']) < TOL
    assert rel_err(c.grad, g['dcoeff']) < TOL and rel_err(m.bias.grad, g['dbias']) < TOL


def test_golden_attention(cuda):
    from feta_tmlr_b200 import DiffTransformerEncoderLayer
    g = _load("attention_case", cuda)
    m = DiffTransformerEncoderLayer(g['d'], g['H'], 2 * g['d'], 0.0).to(cuda)
    m.load_state_dict(g['state_dict'])
    src = g['src'].to(cuda).requires_grad_()
    out, attn, heads = m(src, pe=g['pe'].to(cuda), degree=g['degree'].to(cuda),
                         src_key_padding_mask=g['mask'].to(cuda), need_heads=True)
    ((out * g['w'].to(cuda)).sum() + (heads * g['wh'].to(cuda)).sum()).backward()
    assert rel_err(out, g['out']) < TOL and rel_err(attn, g['attn']) < TOL and rel_err(heads, g['heads']) < TOL
    assert rel_err(src.grad, g['dsrc']) < TOL


def test_golden_model(cuda):
    import feta_tmlr_b200.models as fmodels
    from feta_tmlr_b200 import synthetic
    g = _load("model_case", cuda)
    m = synthetic.build_model("MUTAG", fmodels, **g['over']).to(cuda)
    m.load_state_dict(g['state_dict'])
    b = [None if t is None else t.to(cuda) for t in g['batch']]
    out, _, coeff = m(b[0], b[6], b[7], b[8], b[1], b[2], b[3], b[4], return_filter_coeff=True)
    loss = torch.nn.functional.cross_entropy(out, b[5].long())
    loss.backward()
    assert rel_err(out, g['out']) < TOL and rel_err(coeff, g['coeff']) < TOL
    assert abs(float(loss.detach()) - float(g['loss'])) < 1e-5
    for k, p in m.named_parameters():
        if k in g['grads'] and float(g['grads'][k].abs().max()) > 1e-7:
            assert rel_err(p.grad, g['grads'][k]) < 1e-3, k
