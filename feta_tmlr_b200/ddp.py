"""Data-parallel plumbing: one process per GPU, graphs sharded across
ranks, ONE gradient all-reduce per step.

The reference has no multi-process path (only an ``nn.DataParallel`` flag in non-config scripts,
experiments/run_transformer_gengcn_molpcba.py:446-452).  The path shards naturally -- graphs are
independent -- and has exactly one exchange step, the gradient mean.  All parameter gradients
live as views into one flat fp32 buffer, so the exchange is a single NCCL all-reduce (NVLink 5 /
NVSwitch; ~2.6 MB for the ZINC model, latency-bound) with no per-parameter launches and no
flatten/unflatten copies.
"""
import torch
import torch.distributed as dist


class FlatGradBucket(object):
    """Makes every ``p.grad`` a view into one contiguous buffer and averages it across ranks."""

    def __init__(self, params, process_group=None, attach=True):
        """``attach=False``: only build the flat buffer and its per-parameter views (``self.views``);
        the caller copies gradients in with one multi-tensor copy (engine.GraphedTrainStep)."""
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        total = sum(p.numel() for p in self.params)
        dev, dtype = self.params[0].device, self.params[0].dtype
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        off = 0
        self.views = []
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            if attach:
                p.grad = self.views[-1]                      # autograd accumulates in place
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(dist.get_world_size(self.group))

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


def broadcast_parameters(module, src=0, process_group=None):
    """Replicas start from rank 0's weights (same contract as DDP's constructor)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=process_group)


def shard_indices(num_items, rank, world_size):
    """Contiguous, near-equal shard of ``range(num_items)`` for ``rank`` (graphs are independent,
    so any partition is valid; no data-path collective)."""
    base, rem = divmod(num_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_predictions(*tensors, process_group=None):
    """Epoch-end exchange for the metrics the reference computes over the WHOLE evaluation set on one process --
    ``accuracy_SBM``'s confusion matrix (experiments/run_transformer_gengcn_SBM_cv.py:126-143, called per batch at
    :209/:255) and the OGB ``Evaluator`` ROC-AUC over all predictions (run_transformer_gengcn_molhiv.py:215-219):
    every rank contributes the rows its shard produced (row counts may differ), every rank receives the
    concatenation in rank order.  One ``all_gather`` of the row counts and one padded ``all_gather`` per tensor
    (NCCL has no variable-size gather); not on the training path.  Single process: returns the inputs."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return tensors if len(tensors) > 1 else tensors[0]
    world = dist.get_world_size(process_group)
    n = tensors[0].shape[0]
    for t in tensors:
        if t.shape[0] != n:
            raise ValueError("gather_predictions: tensors must share their first dimension")
    dev = tensors[0].device
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev), group=process_group)
    counts = [int(c) for c in counts]
    cap = max(counts)
    out = []
    for t in tensors:
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:n] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=process_group)
        out.append(torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0))
    return tuple(out) if len(out) > 1 else out[0]
