"""Whole-step CUDA-graph execution (B200-first: CUDA streams and graphs, no tracing compiler).

At the BASELINE shapes a training step is several hundred tiny kernels: the step
is launch-bound, not bandwidth- or math-bound.  ``GraphedTrainStep`` captures
    H2D-staged inputs -> forward_static -> loss -> backward -> gradient all-reduce -> Adam
once, over static shapes (B, Nmax_cap, E_cap), and replays it for every mini-batch: one
``cudaGraphLaunch`` per step instead of ~600 launches.  Mini-batches come from
``data.collate_host(..., static=(nmax_cap, e_cap))``.
"""
import torch

from . import _lib, ddp


def static_caps(store, batch_size, nmax_cap=None, slack=1.15):
    """Static capacities for a dataset: widest graph, and an edge budget per mini-batch."""
    import numpy as np
    lens = store.node_ptr[1:] - store.node_ptr[:-1]
    elens = store.edge_ptr[1:] - store.edge_ptr[:-1]
    nmax = int(lens.max()) if nmax_cap is None else int(nmax_cap)
    top = np.sort(elens)[::-1][:batch_size].sum()                  # worst possible batch
    e_cap = int(min(top, np.ceil(elens.mean() * batch_size * slack) + 4 * elens.max()))
    return nmax, (e_cap + 63) // 64 * 64


class GraphedTrainStep(object):
    """``step(host_batch)`` copies the batch into static device buffers and replays the graph."""

    def __init__(self, model, loss_fn, example_batch, lr=1e-3, device=None, warmup=3):
        self.model = model
        self.loss_fn = loss_fn
        dev = device or next(model.parameters()).device
        self.device = dev
        px, mask, pe, lap, deg, labels, ei = example_batch[:7]
        self.static = [None if t is None else torch.empty_like(t, device=dev) for t in
                       (px, mask, pe, lap, deg, labels, ei)]
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = torch.distributed.get_world_size() if (torch.distributed.is_available() and
                                                           torch.distributed.is_initialized()) else 1
        self.bucket = ddp.FlatGradBucket(self.params, attach=False)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, fused=True, capturable=True)
        self.loss = None
        self.launches_per_step = 0
        self._load(example_batch)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                                # warm-up outside capture
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = _lib.launch_count() - n0           # feta kernels inside one replay

    def _body(self):
        px, mask, pe, lap, deg, labels, ei = self.static
        for p in self.params:
            p.grad = None                       # autograd then WRITES each gradient (no += kernels)
        out = self.model.forward_static(px, ei, mask, pe, lap, deg)
        loss = self.loss_fn(out, labels)
        loss.backward()
        if self.world > 1:                      # one flat all-reduce; the optimizer reads the flat views
            live = [(p, v) for p, v in zip(self.params, self.bucket.views) if p.grad is not None]
            torch._foreach_copy_([v for _, v in live], [p.grad for p, _ in live])
            self.bucket.all_reduce_mean()
            for p, v in live:
                p.grad = v
        self.opt.step()
        self.loss = loss.detach()

    def _load(self, batch):
        for dst, src in zip(self.static, batch[:7]):
            if dst is not None:
                dst.copy_(src, non_blocking=True)

    def step(self, batch=None):
        """``batch``: a static-shape collate tuple (host pinned or device tensors); None = reuse."""
        if batch is not None:
            self._load(batch)
        self.graph.replay()
        return self.loss

    def plan_guard_tripped(self):
        """Synchronising check of the device-side plan guard (see include/feta_b200.h)."""
        plans = list(self.model.encoder.spectral_gnns._plans.values())
        return any(p.meta_host()[7] for p in plans)
