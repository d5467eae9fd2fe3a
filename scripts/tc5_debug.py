import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feta_tmlr_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda")
P = ops._ptr
def fwd(T, fin, fout):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(T, fin, generator=g).to(dev); W = (torch.randn(fout, fin, generator=g) * 0.2).to(dev)
    y = torch.empty(T, fout, device=dev)
    rc = lib.feta_linear_fwd(P(x), P(W), None, P(y), T, fin, fout, 0, torch.cuda.current_stream().cuda_stream)
    ref = x.double() @ W.double().t()
    err = (y.double() - ref).abs()
    bad = (err.max(dim=1).values > 1e-3).nonzero().flatten()
    print("fwd T=%d %d->%d rc=%d max err %.2e bad rows %d" % (T, fin, fout, rc, float(err.max()), bad.numel()),
          (bad[:6].tolist(), bad[-3:].tolist(), sorted(set((bad // 128).tolist()))[:12]) if bad.numel() else "")
    if bad.numel():
        badc = (err[bad[0]] > 1e-3).nonzero().flatten()
        print("   first bad row cols", badc[:4].tolist(), badc[-2:].tolist(), badc.numel())
def dx(T, fin, fout):
    g = torch.Generator().manual_seed(2)
    dy = torch.randn(T, fout, generator=g).to(dev); W = (torch.randn(fout, fin, generator=g) * 0.2).to(dev)
    o = torch.empty(T, fin, device=dev)
    rc = lib.feta_linear_dx(P(dy), P(W), None, None, P(o), T, fin, fout, torch.cuda.current_stream().cuda_stream)
    ref = dy.double() @ W.double()
    err = (o.double() - ref).abs()
    bad = (err.max(dim=1).values > 1e-3).nonzero().flatten()
    print("dx  T=%d in=%d out=%d rc=%d max err %.2e bad rows %d" % (T, fin, fout, rc, float(err.max()), bad.numel()),
          (bad[:6].tolist(), bad[-3:].tolist(), sorted(set((bad // 128).tolist()))[:12]) if bad.numel() else "")
    if bad.numel():
        badc = (err[bad[0]] > 1e-3).nonzero().flatten()
        print("   first bad row cols", badc[:4].tolist(), badc[-2:].tolist(), badc.numel())
for _ in range(2):
    dx(12032, 256, 128)
fwd(12032, 128, 256); fwd(20000, 64, 192); fwd(40000, 64, 64); dx(40000, 64, 64); dx(12032, 192, 64); fwd(227328, 64, 192)
dx(12032, 256, 64); dx(12032, 128, 128); dx(6000, 256, 128)
