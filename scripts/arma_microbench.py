"""GPU micro-benchmark of the fused ARMAConvDynamic kernels (csrc/arma.cu) on an HBM-sized batch of molecule-shape
graphs (the Chebyshev sweep's workload) and at a BASELINE shape, timed as CUDA-graph replays; prints achieved GB/s
against the algorithmic bytes of DESIGN.md section 4.
    python scripts/arma_microbench.py [rows]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feta_tmlr_b200 import ops                          # noqa: E402
from scripts.attn_microbench import time_graphed        # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    dev = torch.device("cuda")
    F, K = 16, 4
    g = torch.Generator(device=dev).manual_seed(0)
    G = rows // 25
    sizes = torch.randint(13, 38, (G,), device=dev, generator=g)
    ptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(sizes, 0)
    R = int(ptr[-1])
    batch = torch.repeat_interleave(torch.arange(G, device=dev), sizes)
    # ring + one chord per graph, both directions
    idx = torch.arange(R, device=dev)
    start = ptr[batch]
    nxt = start + (idx - start + 1) % sizes[batch]
    chord = start + (idx - start + 5) % sizes[batch]
    src = torch.cat([idx, nxt, idx, chord])
    dst = torch.cat([nxt, idx, chord, idx])
    ei = torch.stack([src, dst])
    plan = ops.build_cheb_plan(ei, batch, R, G, 2.0, hints={'max_nodes': 37, 'block_diagonal': True}, norm=ops.NORM_GCN)
    nnz = int(ei.shape[1])
    x = torch.randn(R, F, device=dev, generator=g)
    coeff = torch.randn(G, 2 * K, device=dev, generator=g)
    W = torch.randn(K, F, F, device=dev, generator=g) * 0.2
    V = torch.randn(K, F, F, device=dev, generator=g) * 0.2
    bias = torch.zeros(K, F, device=dev)
    go = torch.randn(R, F, device=dev, generator=g)

    def fwd():
        ops.arma_filter(x, coeff, W, V, bias, plan)

    def fwd_bwd():
        xx = x.detach().requires_grad_()
        cc = coeff.detach().requires_grad_()
        out = ops.arma_filter(xx, cc, W, V, bias, plan)
        torch.autograd.grad(out, (xx, cc), go)

    tf = time_graphed(fwd, reps=5, replays=5)
    tfb = time_graphed(fwd_bwd, reps=5, replays=5)
    fwd_bytes = 12 * R * F + 8 * nnz + 4 * (R + 1) + 8 * G * K
    bwd_bytes = 20 * R * F + 12 * R * K * F + 8 * nnz
    print(json.dumps({"rows": R, "graphs": G, "nnz": nnz, "F": F, "K": K,
                      "fwd_us": round(tf, 1), "fwd_GBps": round(fwd_bytes / tf / 1e3, 1),
                      "bwd_us": round(tfb - tf, 1), "bwd_GBps": round(bwd_bytes / (tfb - tf) / 1e3, 1),
                      "fwd_algorithmic_bytes": fwd_bytes, "bwd_algorithmic_bytes": bwd_bytes}))


if __name__ == "__main__":
    main()
