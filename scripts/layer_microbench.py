"""GPU micro-benchmark of the building blocks of one encoder layer at a BASELINE shape, each timed as CUDA-graph
replays (pure device time per launch; an eager loop would time CPU launch gaps instead).
    python scripts/layer_microbench.py ZINC|PATTERN"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feta_tmlr_b200 import ops, synthetic, data as fdata, engine   # noqa: E402
from scripts.attn_microbench import time_graphed                    # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ZINC"
    cfg = synthetic.CONFIGS[name]
    B, H, d = cfg['batch'], cfg['heads'], cfg['d_model']
    graphs = synthetic.make_dataset(name, B, seed=0)
    store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
    caps = engine.static_caps(store, B)
    b = fdata.collate_host(store, np.arange(B), static=caps)
    dev = torch.device("cuda")
    mask, pe = b[1].to(dev), (None if b[2] is None else b[2].to(dev))
    nmax = mask.shape[1]
    T = nmax * B
    ops.WGRAD_SIDE_STREAM = False
    x = torch.randn(nmax, B, d, device=dev)
    W = torch.randn(d, d, device=dev) * 0.1
    W3 = torch.randn(3 * d, d, device=dev) * 0.1
    W1 = torch.randn(2 * d, d, device=dev) * 0.1
    W2 = torch.randn(d, 2 * d, device=dev) * 0.1
    b1 = torch.zeros(2 * d, device=dev)
    bd = torch.zeros(d, device=dev)
    g, be = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    h = torch.randn(nmax, B, 2 * d, device=dev)
    dy3 = torch.randn(nmax, B, 3 * d, device=dev)
    qkv = torch.randn(nmax, B, 3 * d, device=dev)
    go = torch.randn(nmax, B, H, d // H, device=dev)
    scale = (d // H) ** -0.5
    lib = ops._lib.load()
    res = {"config": name, "tokens": T, "nmax": nmax, "B": B}

    def t(key, fn):
        res[key] = round(time_graphed(fn), 2)

    t("in_proj_fwd", lambda: torch.nn.functional.linear(x, W3))
    t("out_proj_fwd", lambda: torch.nn.functional.linear(x, W, bd))
    t("ffn1_relu_fwd", lambda: torch._addmm_activation(b1, x.view(-1, d), W1.t()))
    t("ffn2_fwd", lambda: torch.nn.functional.linear(h, W2, bd))
    t("in_proj_dx_addmm", lambda: torch.addmm(x.view(-1, d), dy3.view(-1, 3 * d), W3))
    t("ffn2_dx", lambda: x.matmul(W2))
    t("ffn1_dx_addmm", lambda: torch.addmm(x.view(-1, d), h.view(-1, 2 * d), W1))
    t("relu_bwd", lambda: torch.ops.aten.threshold_backward(h, h, 0))
    t("add", lambda: x + x)
    t("ln_fwd", lambda: ops.add_layer_norm(x, x, g, be))

    def ln_fb():
        a = x.detach().requires_grad_()
        y = ops.add_layer_norm(a, x, g, be)
        torch.autograd.grad(y, a, x)
    t("ln_fwd_bwd", ln_fb)

    def lin_fb():
        w = W3.detach().requires_grad_()
        y = ops.linear(x, w)
        torch.autograd.grad(y, w, dy3)
    t("in_proj_fwd_wgrad", lin_fb)
    P = ops._ptr
    st = torch.cuda.current_stream().cuda_stream
    y3 = torch.empty(nmax, B, 3 * d, device=dev)
    y1 = torch.empty(nmax, B, d, device=dev)
    y2 = torch.empty(nmax, B, 2 * d, device=dev)
    t("tc_in_proj_fwd", lambda: lib.feta_linear_fwd(P(x), P(W3), None, P(y3), T, d, 3 * d, 0, torch.cuda.current_stream().cuda_stream))
    t("tc_out_proj_fwd", lambda: lib.feta_linear_fwd(P(x), P(W), P(bd), P(y1), T, d, d, 0, torch.cuda.current_stream().cuda_stream))
    t("tc_ffn1_relu_fwd", lambda: lib.feta_linear_fwd(P(x), P(W1), P(b1), P(y2), T, d, 2 * d, 1, torch.cuda.current_stream().cuda_stream))
    t("tc_ffn2_fwd", lambda: lib.feta_linear_fwd(P(h), P(W2), P(bd), P(y1), T, 2 * d, d, 0, torch.cuda.current_stream().cuda_stream))
    t("tc_in_proj_dx", lambda: lib.feta_linear_dx(P(dy3), P(W3), P(x), None, P(y1), T, d, 3 * d, torch.cuda.current_stream().cuda_stream))
    t("tc_ffn2_dx", lambda: lib.feta_linear_dx(P(x), P(W2), None, P(h), P(y2), T, 2 * d, d, torch.cuda.current_stream().cuda_stream))
    t("tc_ffn1_dx", lambda: lib.feta_linear_dx(P(h), P(W1), P(x), None, P(y1), T, d, 2 * d, torch.cuda.current_stream().cuda_stream))
    t("attn_fwd", lambda: ops.diff_attention(qkv, pe, mask, H, scale))

    def attn_fb():
        q = qkv.detach().requires_grad_()
        a, o = ops.diff_attention(q, pe, mask, H, scale)
        torch.autograd.grad(o, q, go)
    t("attn_fwd_bwd", attn_fb)
    res["ln_bwd"] = round(res["ln_fwd_bwd"] - res["ln_fwd"], 2)
    res["in_proj_wgrad"] = round(res["in_proj_fwd_wgrad"] - res["in_proj_fwd"], 2)
    res["attn_bwd"] = round(res["attn_fwd_bwd"] - res["attn_fwd"], 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
