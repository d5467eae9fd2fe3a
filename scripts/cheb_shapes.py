"""Chebyshev filter at the BASELINE shapes: forward / backward device time per launch (CUDA-graph replays of the op
alone on one real mini-batch), for each kernel family.  python scripts/cheb_shapes.py [ZINC PATTERN ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import time_graphed  # noqa: E402
from feta_tmlr_b200 import data as fdata, ops, synthetic  # noqa: E402
import feta_tmlr_b200.models as fmodels  # noqa: E402


def run(name, dev):
    cfg = synthetic.CONFIGS[name]
    B = cfg['batch']
    graphs = synthetic.make_dataset(name, B, seed=0)
    store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
    b = tuple(None if t is None else t.to(dev) for t in fdata.collate_host(store, np.arange(B))[:9])
    m = synthetic.build_model(name, fmodels).to(dev)
    ctx = m.encoder.batch_context(b[6], b[8], b[7], b[1], b[1].shape[1])
    H, dh = cfg['heads'], cfg['d_model'] // cfg['heads']
    R, G = H * b[8].shape[0], H * B
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(R, dh, device=dev, generator=g)
    th = (torch.randn(G, 4 * dh * dh, device=dev, generator=g) * 0.1).reshape(G, 4, dh, dh).permute(1, 0, 2, 3)
    bias = torch.zeros(dh, device=dev)
    go = torch.randn(R, dh, device=dev, generator=g)
    res = {}
    for label, env in (("tile", {}), ("no_tile", {"FETA_CHEB_NO_TILE_KERNEL": "1"}),
                       ("chunk", {"FETA_CHEB_NO_TILE_KERNEL": "1", "FETA_CHEB_NO_WARP_KERNEL": "1"})):
        for k in ("FETA_CHEB_NO_TILE_KERNEL", "FETA_CHEB_NO_WARP_KERNEL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        fwd = time_graphed(lambda: ops.cheb_filter(x, th, bias, ctx.plan), dev)

        def fb():
            xr, tr = x.detach().requires_grad_(), th.detach().requires_grad_()
            torch.autograd.grad(ops.cheb_filter(xr, tr, bias, ctx.plan), (xr, tr), go)
        both = time_graphed(fb, dev)
        res[label] = (round(fwd, 1), round(both - fwd, 1))
    print(name, "R=%d G=%d F=%d nmax=%d nnz=%d" % (R, G, dh, b[1].shape[1], ctx.plan.meta_host()[0]),
          {k: "fwd %.1f us, bwd %.1f us" % v for k, v in res.items()}, flush=True)


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    for name in (sys.argv[1:] or ["ZINC", "PATTERN", "CLUSTER", "MOLHIV"]):
        run(name, dev)
