"""Whole-step CUDA-graph execution (B200-first: CUDA streams and graphs, no tracing compiler).

At the BASELINE shapes a training step is several hundred tiny kernels: the step
is launch-bound, not bandwidth- or math-bound.  ``GraphedTrainStep`` captures
    H2D-staged inputs -> forward_static -> loss -> backward -> gradient all-reduce -> Adam
once, over static shapes (B, Nmax_cap, E_cap), and replays it for every mini-batch: one
``cudaGraphLaunch`` per step instead of ~600 launches.  Mini-batches come from
``data.collate_host(..., static=(nmax_cap, e_cap))``.
"""
import torch

from . import _lib, ddp, ops


def static_caps(store, batch_size, nmax_cap=None, slack=1.15):
    """Static capacities for a dataset: widest graph, and an edge budget per mini-batch."""
    import numpy as np
    lens = store.node_ptr[1:] - store.node_ptr[:-1]
    elens = store.edge_ptr[1:] - store.edge_ptr[:-1]
    nmax = int(lens.max()) if nmax_cap is None else int(nmax_cap)
    top = np.sort(elens)[::-1][:batch_size].sum()                  # worst possible batch
    e_cap = int(min(top, np.ceil(elens.mean() * batch_size * slack) + 4 * elens.max()))
    return nmax, (e_cap + 63) // 64 * 64


class FlatAdam(object):
    """Adam / AdamW (decoupled ``weight_decay``) over flat buffers: parameters are re-pointed to views of one
    flat fp32 tensor laid out like ``bucket.flat`` (the flat gradient buffer), moments are flat too, and a step is
    ``feta_adam_step`` -- one elementwise kernel instead of torch's four multi-tensor launches.  Same update rule
    as ``torch.optim.Adam`` / ``AdamW`` (experiments/run_transformer_gengcn.py:302, ..._SBM_cv.py:371).
    The step counter and ``lr`` are device scalars (CUDA-graph safe; ``set_lr`` needs no re-capture)."""

    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.bucket = bucket
        flat_g = bucket.flat
        self.flat_p = torch.empty_like(flat_g)
        off = 0
        with torch.no_grad():
            for p in bucket.params:
                n = p.numel()
                view = self.flat_p[off:off + n].view_as(p)
                view.copy_(p)
                p.data = view                                  # the module now reads / is updated through the flat buffer
                off += n
        self.exp_avg = torch.zeros_like(flat_g)
        self.exp_avg_sq = torch.zeros_like(flat_g)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=flat_g.device)
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=flat_g.device)
        self.betas, self.eps, self.weight_decay = (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)

    def set_lr(self, lr):
        self.lr.fill_(float(lr))

    def step(self, grad_scale=1.0):
        """Consumes ``bucket.flat`` (gradients; slots of parameters without a gradient must hold zeros)."""
        lib = _lib.load()
        st = torch.cuda.current_stream(self.flat_p.device).cuda_stream
        _lib.check(lib.feta_adam_step(self.flat_p.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.flat_p.numel(), self.lr.data_ptr(),
                                      self.betas[0], self.betas[1], self.eps, self.weight_decay, float(grad_scale),
                                      self.step_count.data_ptr(), st), "feta_adam_step")


class GraphedTrainStep(object):
    """``step(host_batch)`` copies the batch into static device buffers and replays the graph.

    ``double_buffer=True``: two sets of static input buffers, one captured graph per set (shared memory
    pool), used alternately; ``step(batch, prefetch=next_batch)`` then issues the host->device copy of the
    NEXT mini-batch on a copy stream while the current step's graph runs, so the copy leaves the critical
    path of an end-to-end step (it is still paid for every step)."""

    def __init__(self, model, loss_fn, example_batch, lr=1e-3, device=None, warmup=3, double_buffer=False,
                 flat_adam=True, weight_decay=0.0, comm_slices=3):
        self.model = model
        self.loss_fn = loss_fn
        dev = device or next(model.parameters()).device
        self.device = dev
        self.nsets = 2 if double_buffer else 1
        self.sets = [[None if t is None else torch.empty_like(t, device=dev) for t in example_batch[:7]]
                     for _ in range(self.nsets)]
        self.static = self.sets[0]
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = torch.distributed.get_world_size() if (torch.distributed.is_available() and
                                                           torch.distributed.is_initialized()) else 1
        self.bucket = ddp.FlatGradBucket(self.params, attach=False)
        # flat_adam: one-kernel Adam over flat parameter / gradient / moment buffers (FlatAdam);
        # otherwise torch's fused multi-tensor Adam (4 launches per step)
        self.flat_adam = bool(flat_adam)
        if self.flat_adam:
            self.opt = FlatAdam(self.bucket, lr=lr, weight_decay=weight_decay)
        elif weight_decay:
            self.opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay, fused=True,
                                         capturable=True)
        else:
            self.opt = torch.optim.Adam(model.parameters(), lr=lr, fused=True, capturable=True)
        # Gradient exchange overlapped with the backward pass (world > 1, flat Adam): the flat bucket is cut into
        # `comm_slices` contiguous slices in parameter order; gradients arrive last-layer-first, so the LAST slice
        # completes early in the backward pass.  When every gradient of a slice has been produced (per-parameter
        # post-accumulate hooks) the slice is copied into the flat buffer and all-reduced on a communication stream
        # (forked from the main stream and from the weight-gradient side stream) while the rest of the backward
        # pass runs; only the first slice's all-reduce (embedding + first layers) is exposed before Adam.
        import os
        comm_slices = int(os.environ.get("FETA_COMM_SLICES", comm_slices))
        self.comm_slices = int(comm_slices) if (self.world > 1 and self.flat_adam) else 0
        self.comm_stream = torch.cuda.Stream(device=dev) if self.comm_slices else None
        self._slices, self._hooks, self._works = [], [], []
        self.losses = [None] * self.nsets
        self.loss = None
        # guard words of the plans every step rebuilds, accumulated where no graph temporary can alias them: the two
        # captured graphs share one memory pool, so a tensor allocated inside one capture (the plan's meta array) is
        # only defined until the other graph's next replay
        self._guard_acc = torch.zeros(1, dtype=torch.int32, device=dev)
        self.launches_per_step = 0
        self.cur = 0
        self._in_body = False
        self.copy_stream = torch.cuda.Stream(device=dev) if double_buffer else None
        self._staged = None                       # (batch object, its H2D-complete event)
        self._done = [None] * self.nsets          # event after the last replay that read set s
        for s in range(self.nsets):
            self._load(example_batch, s)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                                # warm-up outside capture
            for _ in range(warmup):
                self._body(0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # optimizer steps taken on the example batch before the graphs exist (a training script that must reproduce
        # a single-process run counts them; tests/ddp_two_rank_check.py does)
        self.uncaptured_steps = int(warmup)
        if self.comm_slices:
            self._setup_comm_slices()
            side.wait_stream(torch.cuda.current_stream(dev))
            self.uncaptured_steps += 1
            with torch.cuda.stream(side):                            # one sliced step outside capture (NCCL warm-up)
                self._body(0)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
        self.graphs = []
        for s in range(self.nsets):
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g, pool=self.graphs[0].pool() if s else None):
                self._body(s)
            if s == 0:
                self.launches_per_step = _lib.launch_count() - n0   # feta kernels inside one replay
            self.graphs.append(g)
        self.graph = self.graphs[0]

    # -- overlapped gradient exchange ---------------------------------------------------------------
    def _setup_comm_slices(self):
        """Cut the live parameters (those that received a gradient in the warm-up steps) into contiguous slices of
        roughly equal bytes and hook them."""
        live = [i for i, p in enumerate(self.params) if p.grad is not None]
        total = sum(self.params[i].numel() for i in live)
        bounds, acc, target = [], 0, total / self.comm_slices
        for i in live:
            acc += self.params[i].numel()
            if acc >= target * (len(bounds) + 1) and len(bounds) < self.comm_slices - 1:
                bounds.append(i + 1)
        edges = [0] + bounds + [len(self.params)]
        offs = [0]
        for p in self.params:
            offs.append(offs[-1] + p.numel())
        self._slices = []
        for a, b in zip(edges[:-1], edges[1:]):
            idx = [i for i in live if a <= i < b]
            if idx:
                self._slices.append({"idx": idx, "lo": offs[a], "hi": offs[b], "left": 0})
        for sl in self._slices:
            for i in sl["idx"]:
                self._hooks.append(self.params[i].register_post_accumulate_grad_hook(self._make_hook(sl)))

    def _make_hook(self, sl):
        def hook(_p):
            sl["left"] -= 1
            if sl["left"] == 0 and self._in_body:
                self._reduce_slice(sl)
        return hook

    def _reduce_slice(self, sl):
        dev = self.device
        main, cs = torch.cuda.current_stream(dev), self.comm_stream
        cs.wait_stream(main)
        for st in ops.side_streams_in_use(dev):                    # weight gradients are produced there
            cs.wait_stream(st)
        with torch.cuda.stream(cs):
            torch._foreach_copy_([self.bucket.views[i] for i in sl["idx"]], [self.params[i].grad for i in sl["idx"]])
            self._works.append(torch.distributed.all_reduce(self.bucket.flat[sl["lo"]:sl["hi"]],
                                                            op=torch.distributed.ReduceOp.SUM, async_op=True))

    def _body(self, s=0):
        px, mask, pe, lap, deg, labels, ei = self.sets[s]
        for p in self.params:
            p.grad = None                       # autograd then WRITES each gradient (no += kernels)
        out = self.model.forward_static(px, ei, mask, pe, lap, deg)
        loss = self.loss_fn(out, labels)
        sliced = bool(self._slices)
        if sliced:
            for sl in self._slices:
                sl["left"] = len(sl["idx"])
            self._works = []
        self._in_body = True
        with ops.wgrad_side_stream(True):       # safe here: every p.grad is None and is read only after the pass
            loss.backward()
        self._in_body = False
        if sliced:
            main = torch.cuda.current_stream(self.device)
            for w in self._works:
                w.wait()
            main.wait_stream(self.comm_stream)
            self.opt.step(grad_scale=1.0 / self.world)
        elif self.flat_adam:                    # gradients -> flat buffer (one multi-tensor copy), one SUM all-reduce,
            live = [(p, v) for p, v in zip(self.params, self.bucket.views) if p.grad is not None]   # one Adam kernel
            torch._foreach_copy_([v for _, v in live], [p.grad for p, _ in live])
            if self.world > 1:
                torch.distributed.all_reduce(self.bucket.flat, op=torch.distributed.ReduceOp.SUM)
            self.opt.step(grad_scale=1.0 / self.world)      # the 1/world of the mean is folded into the update
        else:
            if self.world > 1:                  # one flat all-reduce; the optimizer reads the flat views
                live = [(p, v) for p, v in zip(self.params, self.bucket.views) if p.grad is not None]
                torch._foreach_copy_([v for _, v in live], [p.grad for p, _ in live])
                self.bucket.all_reduce_mean()
                for p, v in live:
                    p.grad = v
            self.opt.step()
        sp = getattr(self.model.encoder, '_static_plans', None)
        if sp:
            self._guard_acc.add_(sp[-1].meta[7:8])          # FETA_META_GUARD of the plan this step built
        self.losses[s] = loss.detach()
        self.loss = self.losses[s]

    def _load(self, batch, s=0):
        for dst, src in zip(self.sets[s], batch[:7]):
            if dst is not None:
                dst.copy_(src, non_blocking=True)

    def step(self, batch=None, prefetch=None):
        """``batch``: a static-shape collate tuple (host pinned or device tensors); None = reuse.
        ``prefetch`` (double_buffer only): the batch the NEXT call will be given."""
        s = self.cur
        main = torch.cuda.current_stream(self.device)
        if batch is not None:
            if self._staged is not None and self._staged[0] is batch:
                main.wait_event(self._staged[1])        # its copy was issued during the previous step
            else:
                self._load(batch, s)
        self._staged = None
        self.graphs[s].replay()
        self.loss = self.losses[s]
        if self.nsets == 2:
            ev = torch.cuda.Event()
            ev.record(main)
            self._done[s] = ev
            o = 1 - s
            if prefetch is not None:
                cs = self.copy_stream
                if self._done[o] is not None:
                    cs.wait_event(self._done[o])        # the last replay that read buffer set `o`
                with torch.cuda.stream(cs):
                    self._load(prefetch, o)
                    ev2 = torch.cuda.Event()
                    ev2.record(cs)
                self._staged = (prefetch, ev2)
            self.cur = o
        return self.loss

    def plan_guard_tripped(self):
        """Synchronising check of the device-side plan guard (see include/feta_b200.h)."""
        return bool(int(self._guard_acc.item()))
