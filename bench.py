#!/usr/bin/env python
"""bench.py -- graphs/s (fwd + bwd + optimizer step) of the FeTA spectral hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One "step" = one training step of the BASELINE configs[1] model (ZINC-shape, batch 128 per GPU,
ChebConvDynamic + diffusion PE beta=1, H=8, L=10, d=64) over one synthetic mini-batch.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "graphs/sec fwd+bwd"
UNIT = "graphs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", default="ZINC", choices=["MUTAG", "ZINC", "PATTERN", "CLUSTER", "MOLHIV"])
    ap.add_argument("--pool", type=int, default=0, help="distinct batches in the rotating pool (0 = auto: > L2)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the HBM-sized Chebyshev kernel sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="override graphs per GPU per step")
    ap.add_argument("--quick", action="store_true", help="leg 1 only (profiling runs)")
    ap.add_argument("--eager", action="store_true", help="eager autograd instead of the whole-step CUDA graph")
    ap.add_argument("--sweep-only", action="store_true", help="run only the HBM-sized Chebyshev sweep")
    ap.add_argument("--sweep-f", type=int, default=0, help="feature width for --sweep-only (default: config's dh)")
    ap.add_argument("--sweep-rows", type=int, default=3_000_000)
    ap.add_argument("--repeats", type=int, default=5, help="timed legs of --steps steps each; the median is reported")
    ap.add_argument("--no-extra", action="store_true", help="skip the PATTERN-shape leg (extra.pattern)")
    ap.add_argument("--no-builder", action="store_true", help="skip the GPU-batch-builder end-to-end leg")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def loss_fn_for(name):
    import torch.nn.functional as F
    if name in ("PATTERN", "CLUSTER", "MUTAG"):
        return lambda out, y: F.cross_entropy(out, y.long())
    if name == "ZINC":
        return lambda out, y: F.l1_loss(out, y)                      # run_transformer_gengcn.py:301
    return lambda out, y: F.binary_cross_entropy_with_logits(out.reshape(-1), y.reshape(-1))


def batch_nbytes(batch):
    return sum(t.numel() * t.element_size() for t in batch if t is not None)


class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, "/tmp/feta_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    smax.append(float(c[2]))
                except ValueError:
                    continue
                for n, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def make_pool(name, cfg, B, n_batches, seed, static=False):
    """Host-side pool of distinct mini-batches (reference collate tuple each).  ``static``: pad every
    batch to the dataset-wide (Nmax, E_cap) so one CUDA graph serves all of them."""
    from feta_tmlr_b200 import data as fdata, engine, synthetic
    graphs = synthetic.make_dataset(name, B * n_batches, seed=seed)
    store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
    caps = engine.static_caps(store, B) if static else None
    # static (CUDA-graph) batches ship the edge list as int32: half the host->device bytes of the step
    return [fdata.collate_host(store, np.arange(i * B, (i + 1) * B), static=caps,
                               edge_dtype=np.int32 if static else None)[:9] for i in range(n_batches)]


def call_model(model, b):
    px, mask, pe, lap, deg, labels, ei, bi, fi = b
    return model(px, ei, bi, fi, mask, pe, lap, deg)


def run_reference(args, cfg, B):
    """The reference's algorithm (CPU restatement, oracle/) on all host cores -- bounded sample."""
    import oracle.models as omodels
    from feta_tmlr_b200 import synthetic
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    pool = make_pool(args.config, cfg, B, 2, seed=0)
    torch.manual_seed(0)
    model = synthetic.build_model(args.config, omodels)         # literal all-pairs GCN (models.py:240-287)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    lf = loss_fn_for(args.config)

    def step(i):
        b = pool[i % len(pool)]
        opt.zero_grad()
        loss = lf(call_model(model, b)[0], b[5])
        loss.backward()
        opt.step()
        return float(loss.detach())

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    return val, dt, cores


def time_graphed(fn, dev, reps=20, replays=10):
    """Device time per call of ``fn`` (us): ``reps`` calls captured into one CUDA graph, replayed
    ``replays`` times between two events -- no CPU launch gaps, warm L2 (the in-step situation)."""
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / (reps * replays) * 1e3


def cheb_algorithmic_bytes(R, F, nnz, G, K):
    """x + out + CSR(colidx, vals, rowptr) + Theta + graph_ptr + bias."""
    return 4 * R * F + 4 * R * F + 8 * nnz + 4 * (R + 1) + 4 * G * K * F * F + 4 * (G + 1) + 4 * F


def cheb_bwd_bytes(R, F, nnz, G, K):
    """DESIGN.md section 4, per backward kernel: dOut (or x + dOut) in, dx (or dTheta) out, CSR, Theta."""
    return 8 * R * F + 8 * nnz + 4 * (R + 1) + 4 * G * K * F * F


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from a committed `ncu --set full` capture
    (profiles/ncu_traffic.json: {key: {"bytes": ..., "capture": "<file>", "git": "<build the capture profiled>"}}).
    A capture describes the build it was taken on, so the entry is returned with its provenance, never silently."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, "no profiles/ncu_traffic.json"
    with open(p) as f:
        d = json.load(f)
    e = d.get(key)
    if e is None:
        return None, "no ncu capture for %s" % key
    return int(e["bytes"]), "ncu --set full, %s (build %s)" % (e.get("capture"), e.get("git"))


def make_sweep_inputs(dev, F, K, target_rows):
    from feta_tmlr_b200 import ops
    g = torch.Generator(device=dev).manual_seed(0)
    G = target_rows // 25
    sizes = torch.randint(10, 41, (G,), device=dev, generator=g)
    gp = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    gp[1:] = torch.cumsum(sizes, 0)
    R = int(gp[-1])
    batch = torch.repeat_interleave(torch.arange(G, device=dev), sizes)
    idx = torch.arange(R, device=dev)
    first = gp[:-1][batch]
    chain = idx[idx != first]                                      # bond (i-1, i) inside each molecule
    extra_src = idx[::9]
    span = sizes[batch[extra_src]]
    extra_dst = first[extra_src] + (extra_src - first[extra_src] + 3) % span   # a few ring closures
    s = torch.cat([chain - 1, extra_src])
    t = torch.cat([chain, extra_dst])
    ei = torch.stack([torch.cat([s, t]), torch.cat([t, s])])
    plan = ops.build_cheb_plan(ei, batch, R, G, 2.0)
    nnz = plan.meta_host()[0]
    x = torch.randn(R, F, device=dev, generator=g)
    theta = (torch.randn(G, K * F * F, device=dev, generator=g) * 0.1).reshape(G, K, F, F).permute(1, 0, 2, 3)
    bias = torch.zeros(F, device=dev)
    go = torch.randn(R, F, device=dev, generator=g)
    return plan, x, theta, bias, go, R, G, nnz


def _time_loop(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def cheb_sweep(dev, hbm_gbs, F=16, K=4, target_rows=3_000_000, backward=True):
    """The Chebyshev kernels on an HBM-sized (>> 126 MB L2) batch of molecule-shape graphs: forward, and the two
    backward kernels timed separately (autograd asked for dx only / dTheta only)."""
    from feta_tmlr_b200 import ops
    plan, x, theta, bias, go, R, G, nnz = make_sweep_inputs(dev, F, K, target_rows)
    wl = "molecule-shape, %d graphs (10..40 nodes), %d rows, %d nnz, K=%d, F=%d" % (G, R, nnz, K, F)

    def entry(kernel, ms, nbytes, key):
        ach = nbytes / (ms * 1e-3) / 1e9
        tr, why = ncu_traffic(key)
        return {"kernel": kernel, "workload": wl + " (working set %.2f GB >> L2)" % (nbytes / 1e9), "bound": "hbm",
                "achieved": round(ach, 1), "peak": hbm_gbs, "unit": "GB/s", "frac": round(ach / hbm_gbs, 4),
                "traffic": tr, "traffic_source": why, "ms_per_launch": round(ms, 4), "algorithmic_bytes": int(nbytes)}
    out = {}
    ms = _time_loop(lambda: ops.cheb_filter(x, theta, bias, plan))
    out["fwd"] = entry("cheb_fwd_lane_kernel<%d,2,%d,fwd>" % (F, K), ms, cheb_algorithmic_bytes(R, F, nnz, G, K),
                       "cheb_sweep_fwd_f%d_%d" % (F, target_rows))
    if backward:
        xr = x.detach().requires_grad_()
        y = ops.cheb_filter(xr, theta, bias, plan)
        ms = _time_loop(lambda: torch.autograd.grad(y, xr, go, retain_graph=True))
        out["bwd_dx"] = entry("cheb_fwd_lane_kernel<%d,2,%d,transposed> (dx)" % (F, K), ms, cheb_bwd_bytes(R, F, nnz, G, K),
                              "cheb_sweep_dx_f%d_%d" % (F, target_rows))
        del y, xr
        tr_ = theta.detach().requires_grad_()
        y = ops.cheb_filter(x, tr_, bias, plan)
        ms = _time_loop(lambda: torch.autograd.grad(y, tr_, go, retain_graph=True))
        out["bwd_dtheta"] = entry("cheb_dtheta_lane_kernel<%d,2,%d>" % (F, K), ms, cheb_bwd_bytes(R, F, nnz, G, K),
                                  "cheb_sweep_dtheta_f%d_%d" % (F, target_rows))
    return out


def kernel_times(cfg, B, model, bq, dev, use_graph):
    """This repo's hot kernels alone on one of the step's own batches, as CUDA-graph replays between two events
    (events cannot bracket a kernel inside a replay of the step graph; an eager pass would time the CPU launch
    gaps of such small kernels instead)."""
    from feta_tmlr_b200 import ops
    mask_b, pe_b = bq[1], bq[2]
    nm, H_, d_ = mask_b.shape[1], cfg['heads'], cfg['d_model']
    gq = torch.Generator(device=dev).manual_seed(1)
    qkv_t = torch.randn(nm, B, 3 * d_, device=dev, generator=gq)
    go_t = torch.randn(nm, B, H_, d_ // H_, device=dev, generator=gq)
    sc = float(d_ // H_) ** -0.5
    us = {}

    # the step runs the matrix-free kernels (csrc/attention_rows.cu) in every layer whose attention matrix nobody
    # reads (all but the last under last_layer_filter) and the matrix-writing kernels (csrc/attention.cu) in the rest
    for tag, need in (("attn", False), ("attn_matrix", True)):
        def attn_f():
            ops.diff_attention(qkv_t, pe_b, mask_b, H_, sc, need_attn=need)

        def attn_fb():
            xq = qkv_t.detach().requires_grad_()
            _, o_ = ops.diff_attention(xq, pe_b, mask_b, H_, sc, need_attn=need)
            torch.autograd.grad(o_, xq, go_t)
        us[tag + "_fwd"] = time_graphed(attn_f, dev)
        us[tag + "_bwd"] = time_graphed(attn_fb, dev) - us[tag + "_fwd"]
    # the layer's four projections (forward and input gradient), csrc/linear_simt.cu
    T_ = nm * B
    lin_bytes = 0
    us["linear_fwd_layer"] = us["linear_dx_layer"] = us["linear_wgrad_layer"] = 0.0
    wg_ok = True
    for (fi, fo, relu) in ((d_, 3 * d_, False), (d_, d_, False), (d_, 2 * d_, True), (2 * d_, d_, False)):
        xl = torch.randn(T_, fi, device=dev, generator=gq)
        wl = torch.randn(fo, fi, device=dev, generator=gq) * 0.1
        bl = torch.zeros(fo, device=dev)
        gl = torch.randn(T_, fo, device=dev, generator=gq)
        tf = time_graphed(lambda: ops.linear(xl, wl, bl, relu=relu), dev)

        def lin_fb():
            xr = xl.detach().requires_grad_()
            torch.autograd.grad(ops.linear(xr, wl, bl, relu=relu), xr, gl)
        us["linear_fwd_layer"] += tf
        us["linear_dx_layer"] += time_graphed(lin_fb, dev) - tf
        lin_bytes += 4 * (T_ * fi + fi * fo + T_ * fo)
        # the weight-gradient pair of the same projection (csrc/dense.cu: wgrad_partial_kernel on the tensor cores +
        # slices_reduce4_kernel); in the step it runs on the side streams beside the chain, so it is reported, not
        # ranked with the chain's kernels.  Never allowed to break the line.
        if wg_ok:
            try:
                def lin_wb():
                    w_, b_ = wl.detach().requires_grad_(), bl.detach().requires_grad_()
                    torch.autograd.grad(ops.linear(xl, w_, b_, relu=relu), (w_, b_), gl)
                us["linear_wgrad_layer"] += max(time_graphed(lin_wb, dev) - tf, 0.0)
            except Exception as e:                       # noqa: BLE001 -- informational measurement only
                sys.stderr.write("bench.py: weight-gradient timing skipped (%s)\n" % str(e)[:120])
                us["linear_wgrad_layer"], wg_ok = 0.0, False
    us["linear_bytes_layer"] = float(lin_bytes)
    if use_graph:
        ctx_t = model.encoder.static_context(bq[6], mask_b, nm)
        Rt, Gt, dh_ = H_ * B * nm, H_ * B, d_ // H_
    else:
        ctx_t = model.encoder.batch_context(bq[6], bq[8], bq[7], mask_b, nm)
        Rt, Gt, dh_ = H_ * bq[8].shape[0], H_ * B, d_ // H_
    x_t = torch.randn(Rt, dh_, device=dev, generator=gq)
    th_t = (torch.randn(Gt, 4 * dh_ * dh_, device=dev, generator=gq) * 0.1).reshape(Gt, 4, dh_, dh_).permute(1, 0, 2, 3)
    bias_t = torch.zeros(dh_, device=dev)
    go_c = torch.randn(Rt, dh_, device=dev, generator=gq)
    us["cheb_fwd"] = time_graphed(lambda: ops.cheb_filter(x_t, th_t, bias_t, ctx_t.plan), dev)

    def cheb_fb():
        xr, tr = x_t.detach().requires_grad_(), th_t.detach().requires_grad_()
        torch.autograd.grad(ops.cheb_filter(xr, tr, bias_t, ctx_t.plan), (xr, tr), go_c)
    us["cheb_bwd"] = time_graphed(cheb_fb, dev) - us["cheb_fwd"]
    return us, ctx_t.plan.meta_host()[0], Rt, nm


def run_config(name, args, dev, rank, world, steps, warmup, repeats, with_e2e=True, with_kernels=True,
               with_builder=False):
    """One BASELINE config end to end: value (inputs resident in HBM), e2e (host buffers), kernel micro-times."""
    import torch.distributed as dist
    import feta_tmlr_b200.models as fmodels
    from feta_tmlr_b200 import _lib, ddp, synthetic
    cfg = dict(synthetic.CONFIGS[name])
    B = (args.batch if (args.batch and name == args.config) else cfg['batch'])
    use_graph = not args.eager
    probe = make_pool(name, cfg, B, 1, seed=1000 + rank, static=use_graph)
    per_batch = batch_nbytes(probe[0][:7])
    n_pool = args.pool or int(min(256, max(8, np.ceil(160e6 / per_batch))))
    pool_host = make_pool(name, cfg, B, n_pool, seed=rank, static=use_graph)
    pool_pinned = [tuple(None if t is None else t.pin_memory() for t in b) for b in pool_host]
    pool_dev = [tuple(None if t is None else t.to(dev) for t in b) for b in pool_host]

    torch.manual_seed(0)
    model = synthetic.build_model(name, fmodels).to(dev)
    ddp.broadcast_parameters(model)
    lf = loss_fn_for(name)
    if cfg['head'] == 'node':
        lf = (lambda out, y: torch.nn.functional.cross_entropy(out.reshape(-1, out.shape[-1]), y.reshape(-1),
                                                               ignore_index=-100)) if use_graph else \
            (lambda out, y: torch.nn.functional.cross_entropy(out, y.long()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = None
    if use_graph:
        from feta_tmlr_b200 import engine
        eng = engine.GraphedTrainStep(model, lf, pool_dev[0], lr=1e-3, device=dev, double_buffer=True)
        bucket = eng.bucket
        step = eng.step                                   # copies the batch into static buffers + 1 graph launch
        launches_per_step = eng.launches_per_step
    else:
        bucket = ddp.FlatGradBucket(model.parameters())
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)

        def step(b):
            bucket.zero()
            loss = lf(call_model(model, b)[0], b[5])
            loss.backward()
            bucket.all_reduce_mean()
            opt.step()
            return loss
        launches_per_step = None

    def timed(fn, K, W, offset=0):
        for i in range(W):
            fn(i + offset)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for i in range(K):
            fn(W + i + offset)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), _lib.launch_count() - n0

    def repeated(fn, offset=0):
        """`repeats` timed legs of exactly `steps` steps each (max over ranks per leg); the MEDIAN leg is reported."""
        legs = []
        for r in range(repeats):
            ms, launches = timed(fn, steps, warmup if r == 0 else 1, offset=offset + r * (steps + 1))
            legs.append((ms, launches))
        legs.sort()
        ms, launches = legs[len(legs) // 2]
        per = [world * B * steps / (m * 1e-3) for m, _ in legs]
        return ms, launches, {"repeats": repeats, "median": round(world * B * steps / (ms * 1e-3), 1),
                              "min": round(min(per), 1), "max": round(max(per), 1)}

    res = {"name": name, "B": B, "cfg": cfg, "n_pool": n_pool, "per_batch": per_batch, "bucket_bytes": bucket.nbytes(),
           "use_graph": use_graph}
    # ---- leg 1: inputs resident in HBM
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    ms, launches, spread = repeated(lambda i: step(pool_dev[i % n_pool]))
    res["clocks"] = sampler.stop() if rank == 0 else None
    if use_graph:
        launches = launches_per_step * steps              # kernels replayed from the captured graph
    res.update(ms=ms, launches=int(launches), value=world * B * steps / (ms * 1e-3), spread=spread)
    if args.quick:
        return res, eng
    # ---- leg 2: end to end through the public API with HOST buffers (H2D + D2H inside the timed region)
    d2h = [0.0]
    if with_e2e:
        def e2e_step(i):
            hb = pool_pinned[i % n_pool]
            if use_graph:                                     # H2D straight into the graph's static buffers; the copy
                loss = eng.step(hb, prefetch=pool_pinned[(i + 1) % n_pool])   # of the NEXT batch overlaps this step
            else:
                loss = step(tuple(None if t is None else t.to(dev, non_blocking=True) for t in hb))
            d2h[0] = float(loss.detach().cpu())              # device -> host read of the step's result
        ms_e2e, _, spread_e2e = repeated(e2e_step, offset=7)
        res["e2e"] = {"value": round(world * B * steps / (ms_e2e * 1e-3), 1), "unit": UNIT,
                      "h2d_bytes_per_step": int(np.mean([batch_nbytes(b[:7]) for b in pool_pinned])),
                      "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / steps, 4), "spread": spread_e2e}
        res["last_loss"] = d2h[0]
    if use_graph and eng.plan_guard_tripped():
        raise RuntimeError("device-side plan guard tripped: static capacities do not cover a batch")
    if with_builder and use_graph:
        # N2 end to end: the dataset lives in HBM, only the graph ids of the step cross PCIe, the GPU batch builder
        # (csrc/collate.cu) writes the static batch and the graph replays on it
        from feta_tmlr_b200 import data as fdata, engine
        graphs = synthetic.make_dataset(name, B * min(n_pool, 16), seed=rank)
        store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
        caps = tuple(int(v) for v in (pool_host[0][0].shape[1], pool_host[0][6].shape[1]))
        try:
            builder = fdata.DeviceBatchBuilder(store, dev)
            nb = min(n_pool, 16)

            def built_step(i):
                ids = np.arange((i % nb) * B, (i % nb + 1) * B)
                loss = eng.step(builder.build(ids, static=caps))
                d2h[0] = float(loss.detach().cpu())
            ms_b, _, spread_b = repeated(built_step, offset=3)
            res["e2e_device_builder"] = {"value": round(world * B * steps / (ms_b * 1e-3), 1), "unit": UNIT,
                                         "h2d_bytes_per_step": 8 * (3 * B + 2), "d2h_bytes_per_step": 4,
                                         "ms_per_step": round(ms_b / steps, 4), "spread": spread_b}
            if eng.plan_guard_tripped():
                enc = eng.model.encoder
                metas = [list(p_.meta_host()[:8]) for p_ in getattr(enc, '_static_plans', [])]
                res["e2e_device_builder"] = {"error": "device-side plan guard tripped on a builder batch",
                                             "plan_meta": metas, "caps": list(caps)}
        except ValueError as e:                     # a batch beyond the static capacities of the host pool
            res["e2e_device_builder"] = {"error": str(e)[:120]}
    if with_kernels and rank == 0:
        us, nnz, rows, nm = kernel_times(cfg, B, model, pool_dev[3 % n_pool], dev, use_graph)
        res.update(kern_us=us, cheb_nnz=nnz, cheb_rows=rows, nm=nm)
        H, d = cfg['heads'], cfg['d_model']
        attn_rows = []
        for b in pool_dev[:min(n_pool, 8)]:
            lens = (~b[1]).sum(1).double()
            N, sumsq = int(lens.sum()), float((lens * lens).sum())
            # q,k,v + pe + attn write + O
            attn_rows.append((3 * 4 * N * d + (4 * sumsq if b[2] is not None else 0) + 4 * H * sumsq + 4 * N * d,
                              4 * H * sumsq, N))
        res["attn_bytes"] = float(np.mean([a[0] for a in attn_rows]))
        res["attn_matrix_bytes"] = float(np.mean([a[1] for a in attn_rows]))     # the attention-matrix write alone
        res["attn_tokens"] = float(np.mean([a[2] for a in attn_rows]))
    return res, eng


def teardown(world, engines):
    """Orderly exit under torchrun: drop the captured graphs (they hold the NCCL communicator's kernels), sync,
    destroy the process group.  NCCL's teardown after a graph-captured collective has been seen to block on this
    pool, so the destroy runs under a watchdog; if it does not return the process still exits 0 -- every rank has
    already passed the final barrier and printed."""
    import torch.distributed as dist
    import threading
    for e in engines:
        if e is not None:
            e.graphs = []
            e.graph = None
    del engines
    import gc
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        done = threading.Event()

        def _destroy():
            try:
                dist.destroy_process_group()
            finally:
                done.set()
        th = threading.Thread(target=_destroy, daemon=True)
        th.start()
        if not done.wait(20.0):
            sys.stderr.write("bench.py: destroy_process_group() did not return within 20 s; exiting\n")
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)


def main():
    args = parse_args()
    from feta_tmlr_b200 import synthetic
    cfg = dict(synthetic.CONFIGS[args.config])
    B = args.batch or cfg['batch']
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def workload_of(name, c, b):
        return ("%s-shape synthetic, batch %d/GPU, ChebConvDynamic K=4, H=%d, L=%d, d=%d, pos_enc=%s, "
                "lap_dim=%d, LayerNorm, Adam" % (name, b, c['heads'], c['layers'], c['d_model'], c['pos_enc'],
                                                 c['lap_dim']))
    workload = workload_of(args.config, cfg, B)

    if args.impl == "reference":
        if rank != 0:
            return 0
        val, dt, cores = run_reference(args, cfg, B)
        line = {"impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": workload, "where": "host CPU, %d threads" % cores},
                "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "%d steps of one %d-graph batch each (oracle/, literal reference op "
                                           "sequence incl. all-pairs GCN; pinned against the reference run, "
                                           "tests/test_reference_pin.py)" % (args.steps, B)},
                "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ this repo's CUDA path
    import torch.distributed as dist
    from feta_tmlr_b200 import _lib
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    hbm_gbs, peak_src = peaks()
    if args.sweep_only:
        dh = cfg['d_model'] // cfg['heads']
        print(json.dumps(cheb_sweep(dev, hbm_gbs, F=args.sweep_f or dh, target_rows=args.sweep_rows)))
        return 0

    engines = []
    res, eng = run_config(args.config, args, dev, rank, world, args.steps, args.warmup, args.repeats,
                          with_builder=not args.no_builder)
    engines.append(eng)
    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": round(res["value"], 1),
                              "ms_per_step": round(res["ms"] / args.steps, 4), "quick": True,
                              "spread": res["spread"]}), flush=True)
        teardown(world, engines)
        return 0

    # ---- the north-star scaling shape rides along: PATTERN, batch 64 per GPU (BASELINE configs[2])
    extra = {}
    if not args.no_extra and args.config != "PATTERN":
        pres, peng = run_config("PATTERN", args, dev, rank, world, args.steps, args.warmup, max(3, args.repeats // 2 + 1),
                                with_kernels=True)
        engines.append(peng)
        pc = pres["cfg"]
        extra["pattern"] = {"workload": workload_of("PATTERN", pc, pres["B"]), "value": round(pres["value"], 1),
                            "unit": UNIT, "ms_per_step": round(pres["ms"] / args.steps, 4), "spread": pres["spread"],
                            "e2e": pres.get("e2e"), "gpu_launches": pres["launches"],
                            "allreduce_bytes_per_step": pres["bucket_bytes"] if world > 1 else 0,
                            "kernels_us": {k: round(v, 2) for k, v in pres.get("kern_us", {}).items()
                                           if not k.endswith("_bytes_layer")}}

    line = None
    if rank == 0:
        H, L = cfg['heads'], cfg['layers']
        dh = cfg['d_model'] // H
        kern_us, nm = res["kern_us"], res["nm"]
        attn_bytes = res["attn_bytes"]
        cheb_bytes = float(cheb_algorithmic_bytes(res["cheb_rows"], dh, res["cheb_nnz"], H * B, 4))

        def roof(kernel, nbytes, us_, launches_per_step_, note, key):
            ms_ = us_ * 1e-3
            ach = nbytes / (ms_ * 1e-3) / 1e9
            tr, why = ncu_traffic(key)
            return {"kernel": kernel, "bound": "hbm", "achieved": round(ach, 2), "peak": hbm_gbs, "unit": "GB/s",
                    "frac": round(ach / hbm_gbs, 5), "traffic": tr, "traffic_source": why, "peak_source": peak_src,
                    "algorithmic_bytes": int(nbytes), "us_per_launch": round(ms_ * 1e3, 2),
                    "launches_per_step": launches_per_step_, "note": note}
        small = ("%.2f MB per launch: L2-resident, instruction/latency bound at the BASELINE shape (SURVEY.md F5); "
                 "timed as CUDA-graph replays of the kernel ALONE on one of the step's batches (random q/k/v), not "
                 "inside the step graph")
        from feta_tmlr_b200 import ops as _ops
        rows_on = bool(_ops.attn_rows_enabled(nm, dh))
        # the dominant kernel = the largest (time per launch x launches per step) among this repo's kernels
        lin_us = kern_us.pop("linear_fwd_layer")
        lin_dx_us = kern_us.pop("linear_dx_layer")
        lin_bytes = kern_us.pop("linear_bytes_layer")
        lin_wg_us = kern_us.pop("linear_wgrad_layer", None)
        # static (CUDA-graph) step: every layer runs the matrix-free kernels, the coefficient scalar of the last layer
        # is recomputed from q / k (ops.LazyAttention); eager reference-API forward(): the last layer writes its matrix
        n_rows_layers = (L if res["use_graph"] else L - 1) if rows_on else 0
        rows_bytes = attn_bytes - res["attn_matrix_bytes"]
        cands = {
            "linear_fwd": (lin_us / 4.0, 4 * L, lin_bytes / 4.0, "lsimt::linear_simt_kernel<0> (the layer's four "
                           "projections; average of in_proj / out_proj / linear1 / linear2)"),
            "linear_dx": (lin_dx_us / 4.0, 4 * L, lin_bytes / 4.0, "lsimt::linear_simt_kernel<1> (input gradients of "
                          "the four projections, ReLU mask / residual gradient fused)"),
            "attn_fwd": (kern_us["attn_fwd"], n_rows_layers, rows_bytes, "arows::attn_rows_fwd_kernel<%d>" % dh),
            "attn_bwd": (kern_us["attn_bwd"], n_rows_layers, rows_bytes + 4.0 * res["attn_tokens"] * cfg['d_model'],
                         "arows::attn_rows_bwd_kernel<%d>" % dh),
            "attn_matrix_fwd": (kern_us["attn_matrix_fwd"], L - n_rows_layers, attn_bytes,
                                ("attn_fwd_tiled_kernel<%d>" if (nm > 64 and dh in (8, 16)) else "attn_fwd_kernel<%d>")
                                % dh),
        }
        shares = {k_: v_[0] * v_[1] for k_, v_ in cands.items()}
        top = max(shares, key=shares.get)
        t_us, t_n, t_bytes, t_name = cands[top]
        roofline = roof(t_name, t_bytes, t_us, t_n,
                        "largest (us per launch x launches per step) among this repo's kernels: %s; "
                        % ", ".join("%s %.0f us/step" % (k_, v_) for k_, v_ in sorted(shares.items(), key=lambda kv: -kv[1]))
                        + small % (t_bytes / 1e6), "%s_%s" % (top, args.config))
        roofline["attn_rows_fwd_us_per_launch"] = round(kern_us["attn_fwd"], 2)
        roofline["attn_rows_bwd_us_per_launch"] = round(kern_us["attn_bwd"], 2)
        roofline["attn_matrix_fwd_us_per_launch"] = round(kern_us["attn_matrix_fwd"], 2)
        roofline["attn_matrix_bwd_us_per_launch"] = round(kern_us["attn_matrix_bwd"], 2)
        roofline["linear_fwd_us_per_layer"] = round(lin_us, 2)
        roofline["linear_dx_us_per_layer"] = round(lin_dx_us, 2)
        if lin_wg_us:          # weight-gradient pairs of the four projections (side streams; profiles/r2_launches_final.md)
            roofline["linear_wgrad_us_per_layer"] = round(lin_wg_us, 2)
        roofline_cheb = roof("cheb_fwd_%s_kernel<%d>" % ("lane" if nm <= 64 else "graph", dh), cheb_bytes,
                             kern_us["cheb_fwd"], 1,
                             small % (cheb_bytes / 1e6) + "; roofline_sweep is the same kernel family on an HBM-sized "
                             "batch; rows = %d (padded-domain static layout)" % res["cheb_rows"],
                             "cheb_fwd_%s" % args.config)
        roofline_cheb["cheb_bwd_us_per_launch"] = round(kern_us["cheb_bwd"], 2)
        sweep = None
        if not args.no_sweep:
            sweep = {}
            for F_ in (16, 8):                 # head dims of MUTAG/PATTERN/CLUSTER/molhiv (16) and ZINC (8)
                try:
                    sweep["F%d" % F_] = cheb_sweep(dev, hbm_gbs, F=F_, target_rows=args.sweep_rows)
                except torch.OutOfMemoryError as e:       # bounded sweep; never take the box down
                    sweep["F%d" % F_] = {"error": "OOM: %s" % str(e)[:80]}
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            a2 = argparse.Namespace(**vars(args))
            a2.steps, a2.warmup = 4, 1
            cval, cdt, cores = run_reference(a2, cfg, B)
            cpu_baseline = {"value": round(cval, 2), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "4 timed steps (1 warm-up) of one %d-graph batch each, oracle/ on host cores"
                                      % B}
        use_graph = res["use_graph"]
        line = {"metric": METRIC, "value": round(res["value"], 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(res["ms"] / args.steps, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "graphs_per_step": world * B,
                           "parallelism": "dp%d (graphs sharded; flat gradient bucket of %d B/step all-reduced over NCCL in slices "
                                          "overlapped with the backward pass)"
                                          % (world, res["bucket_bytes"]) if world > 1 else "single GPU",
                           "l2": "inputs rotate through a pool of %d distinct device-resident batches "
                                 "(%.0f MB > 126 MB L2)" % (res["n_pool"], res["n_pool"] * res["per_batch"] / 1e6),
                           "edges": "reference-faithful un-tiled edge_index (SURVEY.md F4)",
                           "timing": "value = median of %d legs of exactly %d steps each (spread: min/max)"
                                     % (args.repeats, args.steps),
                           "execution": ("whole training step (forward, loss, backward, gradient all-reduce, flat Adam) "
                                         "captured as ONE CUDA graph over static shapes (engine.GraphedTrainStep; two "
                                         "graphs over two input-buffer sets so the next batch's H2D copy overlaps), "
                                         "replayed per mini-batch") if use_graph else
                           "eager PyTorch autograd over C-ABI kernels, current stream"},
                "clocks": res["clocks"], "spread": res["spread"],
                "e2e": res.get("e2e"),
                "gpu_launches": res["launches"], "roofline": roofline, "roofline_cheb_in_step": roofline_cheb,
                "roofline_sweep": sweep, "cpu_baseline": cpu_baseline, "last_loss": res.get("last_loss"),
                "extra": dict(extra, e2e_device_builder=res.get("e2e_device_builder"),
                              kernels_us={k: round(v, 2) for k, v in kern_us.items()})}
    if rank == 0:
        print(json.dumps(line), flush=True)
    teardown(world, engines)
    return 0


if __name__ == "__main__":
    sys.exit(main())
