"""GPU micro-benchmark of the three attention forward kernels and two backward kernels at a BASELINE
shape, timed as CUDA-graph replays (pure device time per launch, no CPU launch gaps).
    python scripts/attn_microbench.py ZINC|PATTERN|MUTAG"""
import os, subprocess, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feta_tmlr_b200 import ops, synthetic, data as fdata

def time_graphed(fn, reps=20, replays=10):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays) * 1e3      # us per launch

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ZINC"
    variant = os.environ.get("VARIANT", "default")
    cfg = synthetic.CONFIGS[name]; B = cfg['batch']; H = cfg['heads']; d = cfg['d_model']
    graphs = synthetic.make_dataset(name, B, seed=0)
    store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
    b = fdata.collate_host(store, np.arange(B))
    dev = torch.device("cuda")
    mask, pe = b[1].to(dev), (None if b[2] is None else b[2].to(dev))
    nmax = mask.shape[1]
    qkv = torch.randn(nmax, B, 3 * d, device=dev, requires_grad=True)
    scale = (d // H) ** -0.5
    go = torch.randn(nmax, B, H, d // H, device=dev)
    out = {"config": name, "variant": variant, "B": B, "H": H, "nmax": nmax, "dh": d // H}
    for tag, need in (("", True), ("rows_", False)):      # matrix-writing kernels / matrix-free kernels
        def fwd(): ops.diff_attention(qkv.detach(), pe, mask, H, scale, need_attn=need)
        def fwd_bwd():                     # forward + backward in one captured region (same stream)
            x = qkv.detach().requires_grad_()
            a, o = ops.diff_attention(x, pe, mask, H, scale, need_attn=need)
            torch.autograd.grad(o, x, go)
        f, fb = time_graphed(fwd), time_graphed(fwd_bwd)
        out[tag + "fwd_us"] = round(f, 2)
        out[tag + "bwd_us"] = round(fb - f, 2)
    print(json.dumps(out))

if __name__ == "__main__":
    main()
