"""PyTorch-facing operators over the C ABI of libfeta_b200.so.

PyTorch owns memory, streams and autograd bookkeeping; every computation is a call through
``include/feta_b200.h`` with raw device pointers on the current CUDA stream.  There is no CPU
path: CPU tensors raise.  Reference call sites are cited per op.
"""
import torch

from . import _lib
from ._lib import check

META_NNZ, META_NUM_GRAPHS, META_SORTED, META_BLOCKDIAG, META_MAX_NODES, META_MAX_DEG, \
    META_BAD_INDEX, META_GUARD, META_WORDS = 0, 1, 2, 3, 4, 5, 6, 7, 8

_DT = {torch.int64: 0, torch.int32: 1, torch.float32: 2, torch.float64: 3}

import os as _os
# FETA_ATTN_TC=1 routes the attention forward through the tcgen05 / TMEM kernel (3xTF32 QK^T and PV,
# csrc/attention_tc.cu).  Default off: at FeTA's shapes (dh 8..16, <= 190 nodes) the two contractions
# are < 2 % of the kernel's work and the fp32 CUDA-core kernel is 1.6-2.3x faster.
ATTN_TENSOR_CORES = _os.environ.get("FETA_ATTN_TC", "0") == "1"


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Optional per-kernel CUDA-event timers (bench.py's live roofline measurement): when a name is
# enabled, every launch of that kernel is bracketed by two events on the launching stream.
_TIMERS = {}


def enable_kernel_timer(name):
    _TIMERS[name] = []


def disable_kernel_timers():
    _TIMERS.clear()


def kernel_timer_ms(name):
    """Synchronises; returns the list of per-launch durations (ms) recorded for ``name``."""
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in _TIMERS.get(name, [])]


class _timed(object):
    __slots__ = ("name", "a")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if self.name in _TIMERS:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.name in _TIMERS:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _TIMERS[self.name].append((self.a, b))
        return False


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("feta_tmlr_b200 ops run on CUDA tensors only (sm_100a); there is no "
                               "CPU fallback -- got a %s tensor" % t.device)


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError("feta_tmlr_b200 kernels are fp32; got %s" % t.dtype)
    return t.contiguous()


# =====================================================================================
# A2: Laplacian CSR plan
# =====================================================================================
class ChebPlan(object):
    """Device-resident CSR of L_hat (+ its transpose) and the graph segmentation of one
    mini-batch -- what ``ChebConvDynamic.__norm__`` (ChebNetDynamic.py:108-130) and the
    ``torch.unique`` of :148 recompute on every call in the reference."""

    __slots__ = ("rowptr", "colidx", "vals", "rowptr_t", "colidx_t", "vals_t", "graph_ptr",
                 "row_graph", "meta", "num_rows", "num_edges", "num_graphs", "max_nodes",
                 "block_diagonal", "validated", "_keep")

    def meta_host(self):
        """Synchronises; returns the 8 meta words as a python list."""
        return self.meta.cpu().tolist()

    def validate(self):
        """Synchronising check of the device-side flags; raises like the reference would
        (or where the reference would silently mis-assign filters)."""
        m = self.meta_host()
        if m[META_BAD_INDEX]:
            raise IndexError("ChebConvDynamic: edge_index holds a node id outside [0, %d)" % self.num_rows)
        if not m[META_SORTED]:
            raise ValueError("ChebConvDynamic: `batch` must be sorted (nodes of a graph contiguous)")
        if self.num_rows > 0 and m[META_NUM_GRAPHS] != self.num_graphs:
            raise RuntimeError(
                "ChebConvDynamic: `batch` holds %d graphs but filter_coeff has %d "
                "(repeat_interleave size mismatch, ChebNetDynamic.py:149)" % (m[META_NUM_GRAPHS], self.num_graphs))
        if m[META_GUARD]:
            raise RuntimeError("ChebConvDynamic: a fused kernel refused to run because the host-side "
                               "plan hints (max_nodes=%d, block_diagonal) did not hold: meta=%s"
                               % (self.max_nodes, m))
        self.validated = True
        return m


NORM_CHEB_SYM, NORM_GCN = 0, 1      # FETA_NORM_* of include/feta_b200.h


def build_cheb_plan(edge_index, batch, num_rows, num_graphs, lambda_max=2.0, hints=None, norm=NORM_CHEB_SYM):
    """Build the plan on the current stream.

    ``norm`` -- NORM_CHEB_SYM: the scaled Laplacian of ``ChebConvDynamic.__norm__``; NORM_GCN: the
    ``gcn_norm(add_self_loops=False)`` operator of ``ARMAConvDynamic.forward`` (ChebNetDynamic.py:301-305).

    ``hints`` -- optional dict ``{'max_nodes': int, 'block_diagonal': bool}`` supplied by a caller
    that already knows them on the host (e.g. the padded width Nmax of the mini-batch).  With
    hints no device->host synchronisation happens; the fused kernels verify the hints on the
    device (FETA_META_GUARD).  Without hints the meta words are read back once (one sync),
    which is still one sync fewer than the reference's ``.cpu()`` at models.py:246.
    """
    _need_cuda(edge_index, batch)
    lib = _lib.load()
    dev = edge_index.device
    if edge_index.dtype != torch.int64:
        edge_index = edge_index.long()
    edge_index = edge_index.contiguous()
    E = int(edge_index.shape[1])
    R, G = int(num_rows), int(num_graphs)
    if batch is not None:
        if batch.dtype not in _DT:
            raise TypeError("batch dtype %s not supported" % batch.dtype)
        batch = batch.contiguous()
        if batch.numel() != R:
            raise ValueError("batch has %d entries for %d rows" % (batch.numel(), R))
    p = ChebPlan()
    i32 = dict(dtype=torch.int32, device=dev)
    p.rowptr = torch.empty(R + 1, **i32)
    p.rowptr_t = torch.empty(R + 1, **i32)
    p.colidx = torch.empty(max(E, 1), **i32)
    p.colidx_t = torch.empty(max(E, 1), **i32)
    p.vals = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
    p.vals_t = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
    p.graph_ptr = torch.empty(G + 1, **i32)
    p.row_graph = torch.empty(max(R, 1), **i32)
    p.meta = torch.empty(META_WORDS, **i32)
    ws_bytes = lib.feta_cheb_plan_workspace_bytes(R, E)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.feta_graph_plan_build(
        _ptr(edge_index), E, _ptr(batch), _DT[batch.dtype] if batch is not None else 0, R, G,
        int(norm), float(lambda_max), _ptr(p.rowptr), _ptr(p.colidx), _ptr(p.vals), _ptr(p.rowptr_t),
        _ptr(p.colidx_t), _ptr(p.vals_t), _ptr(p.graph_ptr), _ptr(p.row_graph), _ptr(p.meta),
        _ptr(ws), ws_bytes, _stream()), "feta_graph_plan_build")
    p.num_rows, p.num_edges, p.num_graphs = R, E, G
    p._keep = None
    p.validated = False
    if hints is not None:
        p.max_nodes = int(hints.get("max_nodes", 0))
        p.block_diagonal = bool(hints.get("block_diagonal", True))
    else:
        p.max_nodes, p.block_diagonal = 0, True
        m = p.validate()
        p.max_nodes = int(m[META_MAX_NODES])
        p.block_diagonal = bool(m[META_BLOCKDIAG])
    return p


# =====================================================================================
# A1/A3: fused Chebyshev filter
# =====================================================================================
def _theta_layout(theta):
    """theta [K, G, Fin, Fout] with a contiguous inner [Fin, Fout] block -> (tensor, sk, sg)."""
    K, G, fi, fo = theta.shape
    ok = theta.stride(3) == 1 and theta.stride(2) == fo and theta.stride(0) % 4 == 0 \
        and theta.stride(1) % 4 == 0 and theta.data_ptr() % 16 == 0
    if not ok:
        theta = theta.contiguous()
    return theta, theta.stride(0), theta.stride(1)


class ChebFilterFn(torch.autograd.Function):
    """out = sum_k T_k(x) . theta[k, g(row)] + bias   (ChebNetDynamic.py:162-187)."""

    @staticmethod
    def forward(ctx, x, theta, bias, plan):
        _need_cuda(x, theta, bias)
        lib = _lib.load()
        x = _f32c(x)
        if theta.dtype != torch.float32:
            raise TypeError("filter_coeff must be fp32")
        theta, sk, sg = _theta_layout(theta)
        K, G, fin, fout = theta.shape
        R = x.shape[0]
        if x.shape[1] != fin:
            raise ValueError("x has %d channels, filter expects %d" % (x.shape[1], fin))
        if G != plan.num_graphs or R != plan.num_rows:
            raise ValueError("plan was built for R=%d G=%d, got R=%d G=%d"
                             % (plan.num_rows, plan.num_graphs, R, G))
        out = torch.empty((R, fout), dtype=torch.float32, device=x.device)
        ws_bytes = lib.feta_cheb_workspace_bytes(R, fin, fout, K)
        fused = fin == fout and fin in (4, 8, 16, 32) and plan.block_diagonal
        ws = None if fused else torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        biasc = None if bias is None else _f32c(bias)
        with _timed("cheb_fwd"):
            rc = lib.feta_cheb_fwd(_ptr(x), _ptr(plan.rowptr), _ptr(plan.colidx), _ptr(plan.vals),
                                   _ptr(plan.graph_ptr), _ptr(plan.row_graph), _ptr(plan.meta), _ptr(theta),
                                   sk, sg, _ptr(biasc), _ptr(out), R, G, K, fin, fout, plan.max_nodes,
                                   int(plan.block_diagonal), _ptr(ws), ws_bytes if ws is not None else 0,
                                   _stream())
        if rc == -3 and ws is None:     # graph too large for the fused tile: un-fused path
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            rc = lib.feta_cheb_fwd(_ptr(x), _ptr(plan.rowptr), _ptr(plan.colidx), _ptr(plan.vals),
                                   _ptr(plan.graph_ptr), _ptr(plan.row_graph), _ptr(plan.meta),
                                   _ptr(theta), sk, sg, _ptr(biasc), _ptr(out), R, G, K, fin, fout,
                                   plan.max_nodes, int(plan.block_diagonal), _ptr(ws), ws_bytes, _stream())
        check(rc, "feta_cheb_fwd")
        ctx.save_for_backward(x, theta)
        ctx.plan = plan
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, theta = ctx.saved_tensors
        plan = ctx.plan
        K, G, fin, fout = theta.shape
        R = x.shape[0]
        dout = _f32c(dout)
        need_x, need_t, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], \
            ctx.needs_input_grad[2] and ctx.has_bias
        dx = torch.empty_like(x) if need_x else None
        dtheta = torch.empty_strided(theta.shape, theta.stride(), dtype=torch.float32,
                                     device=x.device) if need_t else None
        dbias = torch.empty(fout, dtype=torch.float32, device=x.device) if need_b else None
        ws_bytes = lib.feta_cheb_workspace_bytes(R, fin, fout, K)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        check(lib.feta_cheb_bwd(_ptr(dout), _ptr(x), _ptr(plan.rowptr), _ptr(plan.colidx), _ptr(plan.vals),
                                _ptr(plan.rowptr_t), _ptr(plan.colidx_t), _ptr(plan.vals_t),
                                _ptr(plan.graph_ptr), _ptr(plan.row_graph), _ptr(plan.meta), _ptr(theta),
                                theta.stride(0), theta.stride(1), _ptr(dx), _ptr(dtheta), _ptr(dbias), R, G,
                                K, fin, fout, plan.max_nodes, int(plan.block_diagonal), _ptr(ws), ws_bytes,
                                _stream()), "feta_cheb_bwd")
        return dx, dtheta, dbias, None


def cheb_filter(x, theta, bias, plan):
    return ChebFilterFn.apply(x, theta, bias, plan)


# =====================================================================================
# N4: fused ARMA filter (ARMAConvDynamic, num_layers = 1)
# =====================================================================================
def _wgrad(dy, x):
    """dW[out, in] = dy^T x and db[out] = colsum(dy) through feta_linear_wgrad."""
    lib = _lib.load()
    T, out_f = dy.shape
    in_f = x.shape[1]
    S = lib.feta_linear_wgrad_slices(T)
    n_part = S * (out_f * in_f + out_f)
    partial = torch.empty(n_part, dtype=torch.float32, device=dy.device)
    dw = torch.empty((out_f, in_f), dtype=torch.float32, device=dy.device)
    db = torch.empty(out_f, dtype=torch.float32, device=dy.device)
    check(lib.feta_linear_wgrad(_ptr(dy), _ptr(x), _ptr(dw), _ptr(db), _ptr(partial), n_part,
                                _ptr(_counters(dy.device)), T, out_f, in_f, _stream()), "feta_linear_wgrad")
    return dw, db


class ArmaFilterFn(torch.autograd.Function):
    """out = 1/K sum_k relu(a[g,k] (A_hat x) W_k + b[g,k] x_root V_k + bias_k)  (ChebNetDynamic.py:297-346)."""

    @staticmethod
    def forward(ctx, x, x_root, coeff, init_weight, root_weight, bias, plan):
        _need_cuda(x, coeff, init_weight, root_weight, bias)
        lib = _lib.load()
        x = _f32c(x)
        same_root = x_root is None
        xr = x if same_root else _f32c(x_root)
        coeff = _f32c(coeff)
        W, V = _f32c(init_weight), _f32c(root_weight)
        b = None if bias is None else _f32c(bias)
        K, F = W.shape[0], W.shape[1]
        R, G = x.shape[0], coeff.shape[0]
        if W.shape != (K, F, F) or V.shape != (K, F, F) or x.shape[1] != F:
            raise ValueError("ARMAConvDynamic needs in_channels == out_channels (ChebNetDynamic.py:284); "
                             "got x %s, init_weight %s, root_weight %s" % (tuple(x.shape), tuple(W.shape), tuple(V.shape)))
        if coeff.shape[1] != 2 * K:
            raise ValueError("filter_coeff must be [G, 2*num_stacks]; got %s" % (tuple(coeff.shape),))
        if G != plan.num_graphs or R != plan.num_rows:
            raise ValueError("plan was built for R=%d G=%d, got R=%d G=%d" % (plan.num_rows, plan.num_graphs, R, G))
        if not plan.block_diagonal:
            raise _lib.FetaError("ARMAConvDynamic: edges must stay inside their graph")
        out = torch.empty((R, F), dtype=torch.float32, device=x.device)
        prop = torch.empty((R, F), dtype=torch.float32, device=x.device)
        check(lib.feta_arma_fwd(_ptr(x), None if same_root else _ptr(xr), _ptr(plan.rowptr), _ptr(plan.colidx),
                                _ptr(plan.vals), _ptr(plan.graph_ptr), _ptr(plan.row_graph), _ptr(plan.meta),
                                _ptr(coeff), _ptr(W), _ptr(V), _ptr(b), _ptr(out), _ptr(prop), R, G, K, F,
                                plan.max_nodes, _stream()), "feta_arma_fwd")
        ctx.save_for_backward(prop, xr, coeff, W, V, b if b is not None else W.new_empty(0))
        ctx.plan, ctx.same_root, ctx.has_bias = plan, same_root, b is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        prop, xr, coeff, W, V, b = ctx.saved_tensors
        plan = ctx.plan
        K, F = W.shape[0], W.shape[1]
        R, G = prop.shape[0], coeff.shape[0]
        dout = _f32c(dout)
        dev = dout.device
        dx = torch.empty((R, F), dtype=torch.float32, device=dev)
        dxr = None if ctx.same_root else torch.empty((R, F), dtype=torch.float32, device=dev)
        dz = torch.empty((R, K * F), dtype=torch.float32, device=dev)
        dza, dzb = torch.empty_like(dz), torch.empty_like(dz)
        dcoeff = torch.zeros((G, 2 * K), dtype=torch.float32, device=dev)
        check(lib.feta_arma_bwd(_ptr(dout), _ptr(prop), _ptr(xr), _ptr(plan.rowptr_t), _ptr(plan.colidx_t),
                                _ptr(plan.vals_t), _ptr(plan.graph_ptr), _ptr(plan.row_graph), _ptr(plan.meta),
                                _ptr(coeff), _ptr(W), _ptr(V), _ptr(b) if ctx.has_bias else None, _ptr(dx),
                                _ptr(dxr), _ptr(dz), _ptr(dza), _ptr(dzb), _ptr(dcoeff), R, G, K, F,
                                plan.max_nodes, _stream()), "feta_arma_bwd")
        # [K*F, F] (row k*F + j, column f)  ->  [K, F(f), F(j)]
        dW = dV = dbias = None
        if ctx.needs_input_grad[3]:
            dW = _wgrad(dza, prop)[0].view(K, F, F).transpose(1, 2).contiguous()
        if ctx.needs_input_grad[4]:
            dV = _wgrad(dzb, xr)[0].view(K, F, F).transpose(1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[5]:
            dbias = _wgrad(dz, xr)[1].view(K, F)
        return dx, dxr, dcoeff, dW, dV, dbias, None


def arma_filter(x, coeff, init_weight, root_weight, bias, plan, x_root=None):
    return ArmaFilterFn.apply(x, x_root, coeff, init_weight, root_weight, bias, plan)


# =====================================================================================
# A6: kernel-biased attention core
# =====================================================================================
def _mask_u8(mask, B, nmax, device):
    if mask is None:
        return torch.zeros((B, nmax), dtype=torch.uint8, device=device)
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return (mask != 0).to(torch.uint8).contiguous()


class DiffAttentionFn(torch.autograd.Function):
    """(attn [B,H,N,N], o [N,B,H,dh] seq-first) = kernel-biased attention of qkv [N,B,3d].

    The attention core of the layer models.py:166-167 calls."""

    @staticmethod
    def forward(ctx, qkv, pe, mask_u8, num_heads, scale, share_qk, drop=None):
        _need_cuda(qkv, pe, mask_u8, drop)
        lib = _lib.load()
        qkv = _f32c(qkv)
        N, B, d3 = qkv.shape
        d = d3 // 3
        H = num_heads
        dh = d // H
        pec = None if pe is None else _f32c(pe)
        if pec is not None and tuple(pec.shape) != (B, N, N):
            raise ValueError("pe must be [B, Nmax, Nmax] = %s, got %s" % ((B, N, N), tuple(pec.shape)))
        attn = torch.empty((B, H, N, N), dtype=torch.float32, device=qkv.device)
        o_heads = torch.empty((N, B, H, dh), dtype=torch.float32, device=qkv.device)   # seq-first
        rowflag = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
        base = qkv.data_ptr()
        qp, kp, vp = base, base + (0 if share_qk else d * 4), base + 2 * d * 4
        attn_out = attn
        if drop is not None:           # attention-weight dropout: `drop` holds 0 or 1/(1-p) per weight
            drop = _f32c(drop)
            if tuple(drop.shape) != (B, H, N, N):
                raise ValueError("drop must be [B, H, Nmax, Nmax]")
            attn_out = torch.empty_like(attn)
            with _timed("attn_fwd"):
                check(lib.feta_attn_fwd_dropout(qp, kp, vp, B * d3, d3, _ptr(pec), _ptr(mask_u8), _ptr(drop), _ptr(attn),
                                                _ptr(attn_out), _ptr(o_heads), B * d, d, _ptr(rowflag), B, H, N, dh,
                                                float(scale), _stream()), "feta_attn_fwd_dropout")
        else:
            with _timed("attn_fwd"):
                check(lib.feta_attn_fwd(qp, kp, vp, B * d3, d3, _ptr(pec), _ptr(mask_u8), _ptr(attn), _ptr(o_heads),
                                        B * d, d, _ptr(rowflag), B, H, N, dh, float(scale), int(ATTN_TENSOR_CORES),
                                        _stream()), "feta_attn_fwd")
        ctx.save_for_backward(qkv, mask_u8, attn, rowflag, drop)
        ctx.cfg = (H, float(scale), bool(share_qk))
        ctx.mark_non_differentiable(rowflag)
        # the coefficient path detaches `attn` (models.py:282): without this autograd would hand backward an
        # attn-sized tensor of zeros per layer (a fill kernel + 4*H*N^2 bytes read by attn_bwd)
        ctx.set_materialize_grads(False)
        return attn_out, o_heads, rowflag

    @staticmethod
    def backward(ctx, d_attn, d_o_heads, _unused):
        lib = _lib.load()
        qkv, mask_u8, attn, rowflag, drop = ctx.saved_tensors
        H, scale, share_qk = ctx.cfg
        N, B, d3 = qkv.shape
        d = d3 // 3
        dh = d // H
        if d_o_heads is None:
            d_o_heads = torch.zeros((N, B, H, dh), dtype=torch.float32, device=qkv.device)
        d_o_heads = _f32c(d_o_heads)
        d_attn_c = None if d_attn is None else _f32c(d_attn)
        dqkv = torch.empty_like(qkv)
        base = qkv.data_ptr()
        qp, kp, vp = base, base + (0 if share_qk else d * 4), base + 2 * d * 4
        db = dqkv.data_ptr()
        with _timed("attn_bwd"):
            if drop is not None:
                check(lib.feta_attn_bwd_dropout(qp, kp, vp, B * d3, d3, _ptr(mask_u8), _ptr(attn), _ptr(rowflag),
                                                _ptr(drop), _ptr(d_o_heads), B * d, d, _ptr(d_attn_c), db, db + d * 4,
                                                db + 2 * d * 4, B * d3, d3, B, H, N, dh, scale, _stream()),
                      "feta_attn_bwd_dropout")
            else:
                check(lib.feta_attn_bwd(qp, kp, vp, B * d3, d3, _ptr(mask_u8), _ptr(attn), _ptr(rowflag),
                                        _ptr(d_o_heads), B * d, d, _ptr(d_attn_c), db, db + d * 4, db + 2 * d * 4,
                                        B * d3, d3, B, H, N, dh, scale, _stream()), "feta_attn_bwd")
        if share_qk:
            dqkv[..., :d] += dqkv[..., d:2 * d]
            dqkv[..., d:2 * d] = 0
        return dqkv, None, None, None, None, None, None


# Layers whose attention matrix nobody reads (models.py:169-173) run the matrix-free kernels of
# csrc/attention_rows.cu; FETA_ATTN_ROWS=0 returns them to the matrix-writing kernels.
ATTN_ROWS = _os.environ.get("FETA_ATTN_ROWS", "1") == "1"


def attn_rows_enabled(nmax, dh):
    return bool(ATTN_ROWS and _lib.load().feta_attn_rows_supported(int(nmax), int(dh)))


class DiffAttentionRowsFn(torch.autograd.Function):
    """o [N,B,H,dh] (seq-first) = the same attention core as DiffAttentionFn without the attention matrix: the
    forward pass keeps (row max, 1/clamped sum, flag) per query row, the backward pass recomputes P.
    For the layers of models.py:166-173 whose matrix is not consumed."""

    @staticmethod
    def forward(ctx, qkv, pe, mask_u8, num_heads, scale, share_qk):
        _need_cuda(qkv, pe, mask_u8)
        lib = _lib.load()
        qkv = _f32c(qkv)
        N, B, d3 = qkv.shape
        d = d3 // 3
        H = num_heads
        dh = d // H
        pec = None if pe is None else _f32c(pe)
        if pec is not None and tuple(pec.shape) != (B, N, N):
            raise ValueError("pe must be [B, Nmax, Nmax] = %s, got %s" % ((B, N, N), tuple(pec.shape)))
        o_heads = torch.empty((N, B, H, dh), dtype=torch.float32, device=qkv.device)   # seq-first
        stats = torch.empty((B, H, N, 4), dtype=torch.float32, device=qkv.device)
        base = qkv.data_ptr()
        qp, kp, vp = base, base + (0 if share_qk else d * 4), base + 2 * d * 4
        with _timed("attn_fwd"):
            check(lib.feta_attn_rows_fwd(qp, kp, vp, B * d3, d3, _ptr(pec), _ptr(mask_u8), _ptr(o_heads), B * d, d,
                                         _ptr(stats), B, H, N, dh, float(scale), _stream()), "feta_attn_rows_fwd")
        ctx.save_for_backward(qkv, pec, mask_u8, stats, o_heads)
        ctx.cfg = (H, float(scale), bool(share_qk))
        ctx.mark_non_differentiable(stats)
        return o_heads, stats

    @staticmethod
    def backward(ctx, d_o_heads, _unused=None):
        lib = _lib.load()
        qkv, pec, mask_u8, stats, o_heads = ctx.saved_tensors
        H, scale, share_qk = ctx.cfg
        N, B, d3 = qkv.shape
        d = d3 // 3
        dh = d // H
        d_o_heads = _f32c(d_o_heads)
        dqkv = torch.empty_like(qkv)
        base = qkv.data_ptr()
        qp, kp, vp = base, base + (0 if share_qk else d * 4), base + 2 * d * 4
        db = dqkv.data_ptr()
        with _timed("attn_bwd"):
            check(lib.feta_attn_rows_bwd(qp, kp, vp, B * d3, d3, _ptr(pec), _ptr(mask_u8), _ptr(stats), _ptr(o_heads),
                                         _ptr(d_o_heads), B * d, d, db, db + d * 4, db + 2 * d * 4, B * d3, d3,
                                         B, H, N, dh, scale, _stream()), "feta_attn_rows_bwd")
        if share_qk:
            dqkv[..., :d] += dqkv[..., d:2 * d]
            dqkv[..., d:2 * d] = 0
        return dqkv, None, None, None, None, None


def dropout_multiplier(shape, p, device, generator=None):
    """0 or 1/(1-p) per attention weight (what ``F.dropout`` multiplies by); capture-safe (torch's Philox state)."""
    keep = torch.rand(shape, device=device, generator=generator) >= p
    return keep.to(torch.float32) * (1.0 / (1.0 - p))


class LazyAttention(object):
    """What a layer hands out in place of its attention matrix when the caller only wants the filter-coefficient
    scalar from it (``need_attn='coeff'``): q/k (detached), the position-encoding kernel, the mask and the row
    statistics of the matrix-free forward pass.  ``coeff_scalar`` recomputes the matrix entries on the fly
    (feta_attn_rows_coeff); ``materialize`` writes the [B, H, Nmax, Nmax] matrix for a caller that wants to look at it."""

    def __init__(self, qkv, pe, mask_u8, stats, num_heads, scale, share_qk):
        self.qkv, self.pe, self.mask_u8, self.stats = qkv.detach(), pe, mask_u8, stats
        self.num_heads, self.scale, self.share_qk = int(num_heads), float(scale), bool(share_qk)

    def coeff_scalar(self, node_ptr, num_nodes, zero_fill=False):
        lib = _lib.load()
        qkv = self.qkv
        N, B, d3 = qkv.shape
        d = d3 // 3
        H = self.num_heads
        alloc = torch.zeros if zero_fill else torch.empty
        s = alloc(H * int(num_nodes), dtype=torch.float32, device=qkv.device)
        base = qkv.data_ptr()
        kp = base + (0 if self.share_qk else d * 4)
        check(lib.feta_attn_rows_coeff(base, kp, B * d3, d3, _ptr(self.pe), _ptr(self.mask_u8), _ptr(self.stats),
                                       _ptr(node_ptr), _ptr(s), B, H, N, d // H, self.scale, int(num_nodes), _stream()),
              "feta_attn_rows_coeff")
        return s

    def materialize(self):
        with torch.no_grad():
            attn, _, _ = DiffAttentionFn.apply(self.qkv, self.pe, self.mask_u8, self.num_heads, self.scale,
                                               self.share_qk, None)
        return attn


def diff_attention(qkv, pe, key_padding_mask, num_heads, scale, share_qk=False, drop=None, need_attn=True):
    """``drop`` (optional, [B, H, Nmax, Nmax]): attention-weight dropout multipliers; the returned attention is the
    dropped one (``F.dropout(P)``), the saved one the un-dropped P.  ``need_attn=False``: the caller does not read
    the attention matrix -- it is returned as None when the matrix-free kernels cover the shape.
    ``need_attn='coeff'``: the caller only needs ``coeff_scalar`` of this layer's matrix -- a ``LazyAttention`` comes
    back in its place (falls back to the matrix when the matrix-free kernels do not cover the shape / dropout)."""
    N, B, d3 = qkv.shape
    mask_u8 = _mask_u8(key_padding_mask, B, N, qkv.device)
    if need_attn is not True and drop is None and qkv.is_cuda and attn_rows_enabled(N, d3 // 3 // num_heads):
        o_sf, stats = DiffAttentionRowsFn.apply(qkv, pe, mask_u8, num_heads, scale, share_qk)
        if need_attn == 'coeff':           # the caller wants the coefficient scalar of this layer, not the matrix
            pec = None if pe is None else _f32c(pe)
            return LazyAttention(_f32c(qkv), pec, mask_u8, stats, num_heads, scale, share_qk), o_sf
        return None, o_sf
    attn, o_sf, _ = DiffAttentionFn.apply(qkv, pe, mask_u8, num_heads, scale, share_qk, drop)
    return attn, o_sf            # o_sf [Nmax, B, H, dh]; out_each_head = o_sf.permute(1, 0, 2, 3)


# =====================================================================================
# A4: filter-coefficient path
# =====================================================================================
def coeff_scalar(attn, key_padding_mask, node_ptr, num_nodes, zero_fill=False):
    """s [H*N]: the per-node scalar the all-ones GCN of models.py:280-282 reduces to.
    Not differentiable (the reference detaches the attention, models.py:282).
    ``zero_fill``: start from zeros (padded layouts, where masked positions own a slot)."""
    if isinstance(attn, LazyAttention):
        return attn.coeff_scalar(node_ptr, num_nodes, zero_fill=zero_fill)
    _need_cuda(attn, node_ptr)
    lib = _lib.load()
    attn = _f32c(attn.detach())
    B, H, nmax, _ = attn.shape
    mask_u8 = _mask_u8(key_padding_mask, B, nmax, attn.device)
    alloc = torch.zeros if zero_fill else torch.empty
    s = alloc(H * int(num_nodes), dtype=torch.float32, device=attn.device)
    check(lib.feta_coeff_scalar(_ptr(attn), _ptr(mask_u8), _ptr(node_ptr), _ptr(s), B, H, nmax,
                                int(num_nodes), _stream()), "feta_coeff_scalar")
    return s


_POOL_BWD_BLOCKS = 148


class CoeffPoolFn(torch.autograd.Function):
    """pooled[g] = mean_{j in [lo_g, hi_g)} tanh(s_j * wbar + gbias)   (models.py:282-283 with x == 1)."""

    @staticmethod
    def forward(ctx, s, seg_lo, seg_hi, wbar, gbias):
        _need_cuda(s, seg_lo, seg_hi, wbar, gbias)
        lib = _lib.load()
        s, wbar, gbias = _f32c(s), _f32c(wbar), _f32c(gbias)
        G = seg_lo.numel()
        C = wbar.numel()
        pooled = torch.empty((G, C), dtype=torch.float32, device=s.device)
        check(lib.feta_coeff_pool_fwd(_ptr(s), _ptr(seg_lo), _ptr(seg_hi), _ptr(wbar), _ptr(gbias), _ptr(pooled),
                                      G, C, _stream()), "feta_coeff_pool_fwd")
        ctx.save_for_backward(s, seg_lo, seg_hi, wbar, gbias)
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        lib = _lib.load()
        s, seg_lo, seg_hi, wbar, gbias = ctx.saved_tensors
        G = seg_lo.numel()
        C = wbar.numel()
        d_pooled = _f32c(d_pooled)
        nblk = max(1, min(_POOL_BWD_BLOCKS, G))
        partial = torch.empty((nblk, 2, C), dtype=torch.float32, device=s.device)
        d_w = torch.empty(C, dtype=torch.float32, device=s.device)
        d_b = torch.empty(C, dtype=torch.float32, device=s.device)
        check(lib.feta_coeff_pool_bwd(_ptr(s), _ptr(seg_lo), _ptr(seg_hi), _ptr(wbar), _ptr(gbias),
                                      _ptr(d_pooled), _ptr(d_w), _ptr(d_b), _ptr(partial), nblk, G, C, _stream()),
              "feta_coeff_pool_bwd")
        return None, None, None, d_w, d_b


def static_context_tensors(masks, edge_index, nmax, num_heads, tile_heads=False):
    """(ei [2, Ecap(*H)] int64 padded-slot ids, slot_ptr [B], seg_lo [H*B], seg_hi [H*B], real [nmax, B, 1]) of a
    static-shape batch through feta_static_context (two launches)."""
    _need_cuda(masks, edge_index)
    lib = _lib.load()
    B = masks.shape[0]
    H = int(num_heads)
    dev = masks.device
    mask_u8 = _mask_u8(masks, B, nmax, dev)
    if edge_index.dtype not in (torch.int32, torch.int64):
        edge_index = edge_index.long()
    edge_index = edge_index.contiguous()
    ecap = edge_index.shape[1]
    heads = H if tile_heads else 1
    ei = torch.empty((2, ecap * heads), dtype=torch.int64, device=dev)
    i32 = torch.empty(3 * B + 2 * H * B, dtype=torch.int32, device=dev)
    node_end, lens, slot_ptr = i32[:B], i32[B:2 * B], i32[2 * B:3 * B]
    seg_lo, seg_hi = i32[3 * B:3 * B + H * B], i32[3 * B + H * B:]
    real = torch.empty((nmax, B, 1), dtype=torch.float32, device=dev)
    check(lib.feta_static_context(_ptr(mask_u8), _ptr(edge_index), 1 if edge_index.dtype == torch.int32 else 0, ecap,
                                  B, int(nmax), H, int(bool(tile_heads)), _ptr(ei), _ptr(node_end), _ptr(lens),
                                  _ptr(slot_ptr), _ptr(seg_lo), _ptr(seg_hi), _ptr(real), _stream()),
          "feta_static_context")
    return ei, slot_ptr, seg_lo, seg_hi, real


class ColSumFn(torch.autograd.Function):
    """``w.sum(dim=0)`` of a 2-D fp32 tensor through feta_colsum (the all-ones GCN's ``colsum(W)``, coeff.cu)."""

    @staticmethod
    def forward(ctx, w):
        _need_cuda(w)
        lib = _lib.load()
        w = _f32c(w)
        R, C = w.shape
        out = torch.empty(C, dtype=torch.float32, device=w.device)
        partial = torch.empty(int(lib.feta_colsum_partial_floats(C)), dtype=torch.float32, device=w.device)
        check(lib.feta_colsum(_ptr(w), R, C, _ptr(out), _ptr(partial), _stream()), "feta_colsum")
        ctx.rows = R
        return out

    @staticmethod
    def backward(ctx, d_out):
        return d_out.unsqueeze(0).expand(ctx.rows, -1)


def colsum(w):
    return ColSumFn.apply(w)


def coeff_pool(s, graph_ptr, wbar, gbias, seg_hi=None):
    """``graph_ptr`` [G+1] (packed plan) or, with ``seg_hi``, explicit [G] segment starts / ends."""
    if seg_hi is None:
        return CoeffPoolFn.apply(s, graph_ptr[:-1], graph_ptr[1:], wbar, gbias)
    return CoeffPoolFn.apply(s, graph_ptr, seg_hi, wbar, gbias)


# =====================================================================================
# A5: pack / unpack / pool
# =====================================================================================
def _fi64(feature_indices):
    if feature_indices.dtype != torch.int64:
        feature_indices = feature_indices.long()
    return feature_indices.contiguous()


class PackHeadsFn(torch.autograd.Function):
    """x[h*N + i] = o_heads[fi[i,0], fi[i,1], h]   (models.py:177-185 + :347)."""

    @staticmethod
    def forward(ctx, o_heads, fi):
        _need_cuda(o_heads, fi)
        lib = _lib.load()
        o_heads = _f32c(o_heads)
        B, nmax, H, dh = o_heads.shape
        N = fi.shape[0]
        x = torch.empty((H * N, dh), dtype=torch.float32, device=o_heads.device)
        check(lib.feta_pack_heads(_ptr(o_heads), _ptr(fi), _ptr(x), N, B, nmax, H, dh, _stream()),
              "feta_pack_heads")
        ctx.save_for_backward(fi)
        ctx.shape = (B, nmax, H, dh)
        return x

    @staticmethod
    def backward(ctx, dx):
        lib = _lib.load()
        (fi,) = ctx.saved_tensors
        B, nmax, H, dh = ctx.shape
        dx = _f32c(dx)
        d_o = torch.empty(ctx.shape, dtype=torch.float32, device=dx.device)
        check(lib.feta_pack_heads_bwd(_ptr(dx), _ptr(fi), _ptr(d_o), fi.shape[0], B, nmax, H, dh, _stream()),
              "feta_pack_heads_bwd")
        return d_o, None


class UnpackHeadsFn(torch.autograd.Function):
    """out[fi[i,1], fi[i,0], h*dh:(h+1)*dh] = y[h*N + i], zeros elsewhere (models.py:200-202)."""

    @staticmethod
    def forward(ctx, y, fi, B, nmax, H):
        _need_cuda(y, fi)
        lib = _lib.load()
        y = _f32c(y)
        dh = y.shape[1]
        N = fi.shape[0]
        out = torch.empty((nmax, B, H * dh), dtype=torch.float32, device=y.device)
        check(lib.feta_unpack_heads(_ptr(y), _ptr(fi), _ptr(out), N, B, nmax, H, dh, _stream()),
              "feta_unpack_heads")
        ctx.save_for_backward(fi)
        ctx.cfg = (B, nmax, H, dh)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        (fi,) = ctx.saved_tensors
        B, nmax, H, dh = ctx.cfg
        d_out = _f32c(d_out)
        N = fi.shape[0]
        dy = torch.empty((H * N, dh), dtype=torch.float32, device=d_out.device)
        check(lib.feta_unpack_heads_bwd(_ptr(d_out), _ptr(fi), _ptr(dy), N, B, nmax, H, dh, _stream()),
              "feta_unpack_heads_bwd")
        return dy, None, None, None, None


def pack_heads(o_heads, feature_indices):
    return PackHeadsFn.apply(o_heads, _fi64(feature_indices))


def unpack_heads(y, feature_indices, B, nmax, H):
    return UnpackHeadsFn.apply(y, _fi64(feature_indices), B, nmax, H)


class SegmentMeanFn(torch.autograd.Function):
    """PyG global_mean_pool over sorted segments (models.py:283)."""

    @staticmethod
    def forward(ctx, x, graph_ptr):
        _need_cuda(x, graph_ptr)
        lib = _lib.load()
        x = _f32c(x)
        G, C = graph_ptr.numel() - 1, x.shape[1]
        out = torch.empty((G, C), dtype=torch.float32, device=x.device)
        check(lib.feta_segment_mean_fwd(_ptr(x), _ptr(graph_ptr), _ptr(out), G, C, _stream()),
              "feta_segment_mean_fwd")
        ctx.save_for_backward(graph_ptr)
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        (graph_ptr,) = ctx.saved_tensors
        d_out = _f32c(d_out)
        G, C = d_out.shape
        dx = torch.zeros((ctx.rows, C), dtype=torch.float32, device=d_out.device)
        check(lib.feta_segment_mean_bwd(_ptr(d_out), _ptr(graph_ptr), _ptr(dx), G, C, _stream()),
              "feta_segment_mean_bwd")
        return dx, None


def segment_mean(x, graph_ptr):
    return SegmentMeanFn.apply(x, graph_ptr)


class MaskedMeanFn(torch.autograd.Function):
    """GlobalAvg1D (models.py:586-595) over x [B, Nmax, C] (any batch/node strides)."""

    @staticmethod
    def forward(ctx, x, mask_u8):
        _need_cuda(x, mask_u8)
        lib = _lib.load()
        if x.dtype != torch.float32:
            raise TypeError("fp32 only")
        if x.stride(2) != 1:
            x = x.contiguous()
        B, nmax, C = x.shape
        out = torch.empty((B, C), dtype=torch.float32, device=x.device)
        check(lib.feta_masked_mean_fwd(_ptr(x), x.stride(0), x.stride(1), _ptr(mask_u8), _ptr(out), B, nmax, C,
                                       _stream()), "feta_masked_mean_fwd")
        ctx.save_for_backward(mask_u8)
        ctx.shape = (B, nmax, C)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        (mask_u8,) = ctx.saved_tensors
        B, nmax, C = ctx.shape
        d_out = _f32c(d_out)
        dx = torch.empty((B, nmax, C), dtype=torch.float32, device=d_out.device)
        check(lib.feta_masked_mean_bwd(_ptr(d_out), _ptr(mask_u8), _ptr(dx), B, nmax, C, _stream()),
              "feta_masked_mean_bwd")
        return dx, None


def masked_mean(x, mask):
    B, nmax, _ = x.shape
    return MaskedMeanFn.apply(x, _mask_u8(mask, B, nmax, x.device))


class GatherRowsFn(torch.autograd.Function):
    """packed[i] = padded[fi[i,0], fi[i,1]]  (models.py:347, :1070-1071); padded [B, Nmax, C] view."""

    @staticmethod
    def forward(ctx, padded, fi):
        _need_cuda(padded, fi)
        lib = _lib.load()
        if padded.dtype != torch.float32:
            raise TypeError("fp32 only")
        if padded.stride(2) != 1:
            padded = padded.contiguous()
        B, nmax, C = padded.shape
        N = fi.shape[0]
        packed = torch.empty((N, C), dtype=torch.float32, device=padded.device)
        check(lib.feta_gather_rows(_ptr(padded), padded.stride(0), padded.stride(1), _ptr(fi), _ptr(packed), N, C,
                                   _stream()), "feta_gather_rows")
        ctx.save_for_backward(fi)
        ctx.shape = (B, nmax, C)
        return packed

    @staticmethod
    def backward(ctx, d_packed):
        lib = _lib.load()
        (fi,) = ctx.saved_tensors
        B, nmax, C = ctx.shape
        d_packed = _f32c(d_packed)
        d_padded = torch.zeros((B, nmax, C), dtype=torch.float32, device=d_packed.device)
        check(lib.feta_scatter_rows(_ptr(d_packed), _ptr(fi), _ptr(d_padded), nmax * C, C, fi.shape[0], C,
                                    _stream()), "feta_scatter_rows")
        return d_padded, None


def gather_rows(padded, feature_indices):
    return GatherRowsFn.apply(padded, _fi64(feature_indices))


# =====================================================================================
# A6 layer glue: token-axis reductions (weight gradients, LayerNorm)
# =====================================================================================
_COUNTERS = {}


def _counters(device):
    """Per-device zero-initialised int32 scratch for the self-re-arming last-block counters of
    csrc/dense.cu (slots 0..255: weight-gradient tiles, slot 256: LayerNorm)."""
    t = _COUNTERS.get(device)
    if t is None:
        t = torch.zeros(512, dtype=torch.int32, device=device)
        _COUNTERS[device] = t
    return t


# Weight gradients are off the critical path of the backward pass (nothing reads them before the
# optimizer / gradient all-reduce), so they CAN run on a per-device side stream: inside a captured step that
# is a parallel branch of the CUDA graph, overlapping the ~70-CTA reduction kernels with the main chain.
# The side stream is joined once per backward pass by an autograd-engine callback.
#
# This is only safe when nothing on the main stream touches those gradients before the join -- autograd
# believes they were produced on the main stream, so an AccumulateGrad that adds into an existing ``p.grad``
# (gradient accumulation, ``zero_grad(set_to_none=False)``, ``FlatGradBucket(attach=True)``) would race with
# the side kernels.  It is therefore OFF unless a caller that guarantees ``p.grad is None`` at backward time
# and consumes the gradients only after the pass opts in with ``wgrad_side_stream(True)``
# (engine.GraphedTrainStep does); FETA_WGRAD_SIDE_STREAM=0 vetoes it everywhere.
_SIDE_ALLOWED = _os.environ.get("FETA_WGRAD_SIDE_STREAM", "1") == "1"
WGRAD_SIDE_STREAM = False


class wgrad_side_stream(object):
    """Context manager / switch: ``with ops.wgrad_side_stream(True): loss.backward()``."""

    def __init__(self, on=True):
        self.on = bool(on) and _SIDE_ALLOWED

    def __enter__(self):
        global WGRAD_SIDE_STREAM
        self.prev = WGRAD_SIDE_STREAM
        WGRAD_SIDE_STREAM = self.on
        return self

    def __exit__(self, *exc):
        global WGRAD_SIDE_STREAM
        WGRAD_SIDE_STREAM = self.prev
        return False


# Opt-in: the layer's projections through csrc/dense_tc.cu (3xTF32 mma.sync GEMMs with fused ReLU-mask /
# residual-gradient epilogues) instead of the library sgemm.  Measured SLOWER on B200 (ZINC shape, in-graph:
# 7.0-13.4 us vs 4.3-7.9 us per GEMM -- three legacy TF32 MMAs per product run at about the SIMT fp32 rate),
# so the default stays the library GEMM; kept, with parity tests, as the base of a tcgen05 version.
LINEAR_TENSOR_CORES = _os.environ.get("FETA_LINEAR_TC", "0") == "1"
# The tcgen05 successor (csrc/linear_tc5.cu: 3xTF32 tcgen05.mma, accumulators in TMEM, bias / ReLU / ReLU-mask /
# residual-gradient epilogues).  FETA_LINEAR_TC5=1 routes every Linear of the layer whose (in, out) are multiples of
# 64 through it (forward and dX); other shapes keep the library GEMM.
LINEAR_TC5 = _os.environ.get("FETA_LINEAR_TC5", "0") == "1"


# Default: the layer's projections (forward and dX) through this repo's fp32 CUDA-core latency kernel
# (csrc/linear_simt.cu) whenever (in, out) are multiples of 64 up to 256 -- every projection of every BASELINE
# config; FETA_LINEAR_SIMT=0 returns them to the library GEMM.
LINEAR_SIMT = _os.environ.get("FETA_LINEAR_SIMT", "1") == "1"
# out_proj -> norm1 and linear2 -> norm2 as ONE launch each (projection with the degree scale, residual add and
# LayerNorm in its epilogue, csrc/linear_simt.cu); FETA_LINEAR_LN_FUSED=0: projection launch + add_layernorm launch
LINEAR_LN_FUSED = _os.environ.get("FETA_LINEAR_LN_FUSED", "1") == "1"
# ... and their backward: LayerNorm backward as the prologue of the projection's dX launch (FETA_LINEAR_LN_BWD_FUSED=0:
# add_layernorm_bwd launch + dX launch)
LINEAR_LN_BWD_FUSED = _os.environ.get("FETA_LINEAR_LN_BWD_FUSED", "1") == "1"
_IMPL_SIMT, _IMPL_TC5, _IMPL_MMA = 1, 2, 3


def linear_impl(in_f, out_f):
    """0: library GEMM; else the FETA_LINEAR_* kernel family that runs this projection (explicit opt-ins win)."""
    lib = _lib.load()
    if LINEAR_TC5 and lib.feta_linear_tc5_supported(int(in_f), int(out_f)):
        return _IMPL_TC5
    if LINEAR_TENSOR_CORES and lib.feta_linear_tc_supported(int(in_f), int(out_f)):
        return _IMPL_MMA
    if LINEAR_SIMT and lib.feta_linear_simt_supported(int(in_f), int(out_f)):
        return _IMPL_SIMT
    return 0


def linear_tc_enabled(in_f, out_f):
    return linear_impl(in_f, out_f) != 0
_SIDE = {}
_JOIN_TASK = {}          # device -> id of the autograd graph task that already queued its join


def _side_stream(device, index=0):
    """index 0: the weight-gradient side stream; 1: the coefficient-branch stream of forward_static."""
    key = device if index == 0 else (device, index)
    st = _SIDE.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE[key] = st
    return st


# Weight-gradient launches of different Linear layers are independent of each other, and the launch list of the
# ZINC step showed them (44 x [partial + reduce] launches) as ONE serial chain that is longer than the main backward
# chain it runs beside (skipping them shortened the step by 17 %): they are dealt round-robin over WGRAD_STREAMS side
# streams, so that chain is cut into that many parallel ones.  Stream 0 is ``_side_stream(device)``.
WGRAD_STREAMS = max(1, int(_os.environ.get("FETA_WGRAD_STREAMS", "3")))
_WG_USED = {}            # device -> side streams that carry work of the running backward pass (join targets)
_WG_NEXT = {}            # device -> round-robin position


def _dev_key(device):
    """One dictionary key per physical device: ``'cuda'`` / ``torch.device('cuda')`` mean the current device."""
    d = torch.device(device)
    if d.type == 'cuda' and d.index is None:
        d = torch.device('cuda', torch.cuda.current_device())
    return d


def _wgrad_stream(device):
    """The side stream the next weight-gradient (or gamma / beta fold) launch goes to; marks it as a join target."""
    device = _dev_key(device)
    i = _WG_NEXT.get(device, 0)
    _WG_NEXT[device] = (i + 1) % WGRAD_STREAMS
    st = _side_stream(device) if i == 0 else _side_stream(device, ('wgrad', i))
    used = _WG_USED.setdefault(device, [])
    if st not in used:
        used.append(st)
    return st


def side_streams_in_use(device):
    """Side streams holding weight-gradient work of the running backward pass (engine: the gradient exchange of a
    slice waits on them)."""
    return list(_WG_USED.get(_dev_key(device), ()))


def _queue_side_join(device):
    """Once per backward pass (keyed by the engine's graph-task id, so a pass that died with an exception cannot
    leave a stale 'already queued' mark): main waits on every side stream the pass used when the pass completes.
    The join is the only place that forgets a stream or rewinds the round-robin, so a nested pass (its own task id,
    its own join) can only join MORE than it launched, never less, and every captured step deals the same streams."""
    task = torch._C._current_graph_task_id()
    if task >= 0 and _JOIN_TASK.get(device) == task:
        return
    _JOIN_TASK[device] = task

    def _join():
        key = _dev_key(device)
        main = torch.cuda.current_stream(device)
        for st in _WG_USED.get(key, ()):
            main.wait_stream(st)
        _WG_USED[key] = []
        _WG_NEXT[key] = 0

    torch.autograd.Variable._execution_engine.queue_callback(_join)


def _linear_wgrad(dy, x, out_f, in_f, has_bias):
    """dW = dy^T x, db = colsum(dy) through feta_linear_wgrad -- on the weight-gradient side stream when a caller
    opted in (wgrad_side_stream), else on the current stream."""
    lib = _lib.load()
    dy2 = _f32c(dy.reshape(-1, out_f))
    x2 = _f32c(x.reshape(-1, in_f))
    T = dy2.shape[0]
    if out_f % 4 or in_f % 4:
        return dy2.t().matmul(x2), (dy2.sum(0) if has_bias else None)
    S = lib.feta_linear_wgrad_slices(T)
    n_part = S * (out_f * in_f + out_f)
    partial = torch.empty(n_part, dtype=torch.float32, device=dy.device)
    dw = torch.empty((out_f, in_f), dtype=torch.float32, device=dy.device)
    db = torch.empty(out_f, dtype=torch.float32, device=dy.device) if has_bias else None
    cnt = _counters(dy.device)
    if WGRAD_SIDE_STREAM:
        main = torch.cuda.current_stream(dy.device)
        side = _wgrad_stream(dy.device)
        side.wait_stream(main)                      # dy, x (and the buffers above) are ready
        check(lib.feta_linear_wgrad(_ptr(dy2), _ptr(x2), _ptr(dw), _ptr(db), _ptr(partial), n_part,
                                    _ptr(cnt), T, out_f, in_f, side.cuda_stream), "feta_linear_wgrad")
        for t in (dy2, x2, partial, dw, db):        # allocator: not reusable until `side` passed here
            if t is not None:
                t.record_stream(side)
        _queue_side_join(dy.device)
    else:
        check(lib.feta_linear_wgrad(_ptr(dy2), _ptr(x2), _ptr(dw), _ptr(db), _ptr(partial), n_part,
                                    _ptr(cnt), T, out_f, in_f, _stream()), "feta_linear_wgrad")
    return dw, db


def _layernorm_backward(dy, z, mean, rstd, gamma, bscale, want_dbs):
    """(dz, dbs, dgamma, dbeta) of y = LayerNorm(z) * gamma + beta with z = a + bscale * b."""
    lib = _lib.load()
    D = z.shape[-1]
    T = z.numel() // D
    dy = _f32c(dy)
    dz = torch.empty_like(z)
    dbs = torch.empty_like(z) if want_dbs else None
    nblk = lib.feta_add_layernorm_bwd_blocks(T)
    partial = torch.empty(nblk * 2 * D, dtype=torch.float32, device=z.device)
    dg = torch.empty(D, dtype=torch.float32, device=z.device)
    db = torch.empty(D, dtype=torch.float32, device=z.device)
    cnt = _counters(z.device)
    if WGRAD_SIDE_STREAM:       # dz on the critical path; the dgamma / dbeta fold on the side stream
        check(lib.feta_add_layernorm_bwd(_ptr(dy), _ptr(z), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(bscale),
                                         _ptr(dz), _ptr(dbs), None, None, _ptr(partial), None, T, D, _stream()),
              "feta_add_layernorm_bwd")
        main = torch.cuda.current_stream(z.device)
        side = _wgrad_stream(z.device)
        side.wait_stream(main)
        check(lib.feta_add_layernorm_bwd_fold(_ptr(partial), T, D, _ptr(dg), _ptr(db), side.cuda_stream),
              "feta_add_layernorm_bwd_fold")
        for t in (partial, dg, db):
            t.record_stream(side)
        _queue_side_join(z.device)
    else:
        check(lib.feta_add_layernorm_bwd(_ptr(dy), _ptr(z), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(bscale),
                                         _ptr(dz), _ptr(dbs), _ptr(dg), _ptr(db), _ptr(partial),
                                         cnt.data_ptr() + 256 * 4, T, D, _stream()), "feta_add_layernorm_bwd")
    return dz, dbs, dg, db


class LinearFn(torch.autograd.Function):
    """y = x W^T + b.  Forward and dX are library GEMMs; dW / db (reductions over the ~5k-token axis
    the libraries under-parallelise here) go through feta_linear_wgrad."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, with_res, mask_input_grad, grad_premasked):
        """``relu``: the activation runs in the GEMM epilogue.  ``with_res``: also return ``x`` itself as a
        second output to be used as the residual input of the following add+LayerNorm, so that backward
        receives the residual gradient and folds it into the dX GEMM instead of leaving a separate
        accumulation kernel to autograd.  ``mask_input_grad``: ``x`` is a ReLU output, so dX is multiplied by
        ``[x > 0]`` in the dX epilogue -- and the layer that produced ``x`` is called with
        ``grad_premasked`` so that it does not apply the ReLU mask again."""
        lib = _lib.load()
        ctx.has_bias = bias is not None
        ctx.relu = bool(relu)
        ctx.mask_in = bool(mask_input_grad)
        ctx.premasked = bool(grad_premasked)
        ctx.set_materialize_grads(False)
        out_f, in_f = weight.shape
        ctx.tc = linear_impl(in_f, out_f) if (x.is_cuda and x.dtype == torch.float32
                                               and weight.dtype == torch.float32) else 0
        if ctx.tc:
            x2 = _f32c(x.reshape(-1, in_f))
            w = _f32c(weight)
            bc = None if bias is None else _f32c(bias)
            y = torch.empty(x.shape[:-1] + (out_f,), dtype=torch.float32, device=x.device)
            check(lib.feta_linear_fwd_ex(_ptr(x2), _ptr(w), _ptr(bc), _ptr(y), x2.shape[0], in_f, out_f,
                                         int(ctx.relu), ctx.tc, _stream()), "feta_linear_fwd")
        elif relu and bias is not None and x.is_cuda:
            y = torch._addmm_activation(bias, x.reshape(-1, in_f), weight.t()).view(*x.shape[:-1], out_f)
        else:
            y = torch.nn.functional.linear(x, weight, bias)
            if relu:
                y = torch.relu_(y)
        ctx.save_for_backward(x, weight, y if (relu and not grad_premasked) else None)
        if with_res:
            return y, x
        return y

    @staticmethod
    def backward(ctx, dy, dres=None):
        lib = _lib.load()
        x, weight, y = ctx.saved_tensors
        out_f, in_f = weight.shape
        dx = dw = db = None
        if dy is None:                                            # only the residual branch was used
            return dres, None, None, None, None, None, None
        if ctx.relu and not ctx.premasked:
            dy = torch.ops.aten.threshold_backward(dy, y, 0)
        if ctx.needs_input_grad[0]:
            if ctx.tc:
                dy2 = _f32c(dy.reshape(-1, out_f))
                x2 = _f32c(x.reshape(-1, in_f)) if ctx.mask_in else None
                dr = None if dres is None else _f32c(dres.reshape(-1, in_f))
                dx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
                check(lib.feta_linear_dx_ex(_ptr(dy2), _ptr(_f32c(weight)), _ptr(dr), _ptr(x2), _ptr(dx),
                                            dy2.shape[0], in_f, out_f, ctx.tc, _stream()), "feta_linear_dx")
            else:
                dx = dy.matmul(weight)
                if ctx.mask_in:
                    dx = dx * (x > 0)
                if dres is not None:
                    dx = dx + dres
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = _linear_wgrad(dy, x, out_f, in_f, ctx.has_bias)
        return dx, dw, db, None, None, None, None


def linear(x, weight, bias=None, relu=False, mask_input_grad=False, grad_premasked=False):
    _need_cuda(x, weight, bias)
    return LinearFn.apply(x, weight, bias, relu, False, mask_input_grad, grad_premasked)


def linear_res(x, weight, bias=None, relu=False, grad_premasked=False):
    """(linear(x), x): the second output is ``x`` for the residual connection (see LinearFn.forward)."""
    _need_cuda(x, weight, bias)
    return LinearFn.apply(x, weight, bias, relu, True, False, grad_premasked)


class AddLayerNormFn(torch.autograd.Function):
    """y = LayerNorm(a + bscale * b) * gamma + beta  (degree scaling + residual + norm of the layer),
    one kernel each way."""

    @staticmethod
    def forward(ctx, a, b, bscale, gamma, beta, eps):
        _need_cuda(a, b, gamma, beta)
        lib = _lib.load()
        a = _f32c(a)
        b = None if b is None else _f32c(b)
        bscale = None if bscale is None else _f32c(bscale)
        D = a.shape[-1]
        T = a.numel() // D
        if bscale is not None and bscale.numel() != T:
            raise ValueError("bscale must have one entry per row")
        y = torch.empty_like(a)
        z = torch.empty_like(a)
        mean = torch.empty(T, dtype=torch.float32, device=a.device)
        rstd = torch.empty(T, dtype=torch.float32, device=a.device)
        gamma, beta = _f32c(gamma), _f32c(beta)
        check(lib.feta_add_layernorm_fwd(_ptr(a), _ptr(b), _ptr(bscale), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(z),
                                         _ptr(mean), _ptr(rstd), T, D, float(eps), _stream()),
              "feta_add_layernorm_fwd")
        ctx.save_for_backward(z, mean, rstd, gamma, bscale)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        z, mean, rstd, gamma, bscale = ctx.saved_tensors
        dz, dbs, dg, db = _layernorm_backward(dy, z, mean, rstd, gamma, bscale, ctx.has_b and bscale is not None)
        grad_b = None
        if ctx.has_b:
            grad_b = dbs if dbs is not None else dz
        return dz, grad_b, None, dg, db, None


def add_layer_norm(a, b, gamma, beta, eps=1e-5, bscale=None):
    return AddLayerNormFn.apply(a, b, bscale, gamma, beta, eps)


class LinearAddLayerNormFn(torch.autograd.Function):
    """y = LayerNorm(res + bscale * (x W^T + b)) * gamma + beta in ONE launch: the projection with the degree scale,
    residual add and LayerNorm in its epilogue (default csrc/linear_simt.cu, exact fp32; FETA_LINEAR_TC5=1:
    csrc/linear_tc5.cu, tcgen05 3xTF32); backward = LayerNorm backward, then the Linear's dX (optional ReLU mask of
    its input) and its weight gradients (side stream when opted in)."""

    @staticmethod
    def forward(ctx, x, weight, bias, res, bscale, gamma, beta, eps, mask_input_grad):
        _need_cuda(x, weight, bias, res, gamma, beta)
        lib = _lib.load()
        out_f, in_f = weight.shape
        x2 = _f32c(x.reshape(-1, in_f))
        res2 = _f32c(res.reshape(-1, out_f))
        w = _f32c(weight)
        bc = None if bias is None else _f32c(bias)
        bscale = None if bscale is None else _f32c(bscale)
        gamma, beta = _f32c(gamma), _f32c(beta)
        T = x2.shape[0]
        y = torch.empty(res.shape, dtype=torch.float32, device=x.device)
        z = torch.empty(res.shape, dtype=torch.float32, device=x.device)
        mean = torch.empty(T, dtype=torch.float32, device=x.device)
        rstd = torch.empty(T, dtype=torch.float32, device=x.device)
        check(lib.feta_linear_layernorm_fwd_ex(_ptr(x2), _ptr(w), _ptr(bc), _ptr(res2), _ptr(bscale), _ptr(gamma),
                                               _ptr(beta), _ptr(y), _ptr(z), _ptr(mean), _ptr(rstd), T, in_f, out_f,
                                               float(eps), linear_layernorm_impl(in_f, out_f), _stream()),
              "feta_linear_layernorm_fwd")
        ctx.save_for_backward(x, w, z, mean, rstd, gamma, bscale)
        ctx.has_bias = bias is not None
        ctx.mask_in = bool(mask_input_grad)
        ctx.fused_bwd = bool(LINEAR_LN_BWD_FUSED and linear_layernorm_impl(in_f, out_f) == _IMPL_SIMT
                             and out_f == 64 and in_f % 64 == 0)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w, z, mean, rstd, gamma, bscale = ctx.saved_tensors
        out_f, in_f = w.shape
        dx = dw = db = None
        if ctx.fused_bwd and ctx.needs_input_grad[0]:
            # LayerNorm backward as the prologue of the projection's dX: one launch (csrc/linear_simt.cu)
            dy = _f32c(dy)
            T = z.numel() // out_f
            dz = torch.empty_like(z)
            dlin = torch.empty_like(z) if bscale is not None else None
            dx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            nblk = lib.feta_lnbwd_linear_dx_blocks(T)
            partial = torch.empty(nblk * 2 * out_f, dtype=torch.float32, device=z.device)
            dg = torch.empty(out_f, dtype=torch.float32, device=z.device)
            dbeta = torch.empty(out_f, dtype=torch.float32, device=z.device)
            x2 = _f32c(x.reshape(-1, in_f)) if ctx.mask_in else None
            check(lib.feta_lnbwd_linear_dx(_ptr(dy), _ptr(z), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(bscale), _ptr(w),
                                           _ptr(x2), _ptr(dz), _ptr(dlin), _ptr(dx), _ptr(partial), T, in_f, out_f,
                                           _stream()), "feta_lnbwd_linear_dx")
            if WGRAD_SIDE_STREAM:                               # the dgamma / dbeta fold leaves the critical path
                main = torch.cuda.current_stream(z.device)
                side = _wgrad_stream(z.device)
                side.wait_stream(main)
                check(lib.feta_ln_fold(_ptr(partial), nblk, out_f, _ptr(dg), _ptr(dbeta), side.cuda_stream),
                      "feta_ln_fold")
                for t in (partial, dg, dbeta):
                    t.record_stream(side)
                _queue_side_join(z.device)
            else:
                check(lib.feta_ln_fold(_ptr(partial), nblk, out_f, _ptr(dg), _ptr(dbeta), _stream()), "feta_ln_fold")
            if dlin is None:
                dlin = dz
            if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
                dw, db = _linear_wgrad(dlin, x, out_f, in_f, ctx.has_bias)
            return dx, dw, db, dz.view(z.shape), None, dg, dbeta, None, None
        dz, dbs, dg, dbeta = _layernorm_backward(dy, z, mean, rstd, gamma, bscale, bscale is not None)
        dlin = dbs if dbs is not None else dz                   # gradient of the Linear's output
        if ctx.needs_input_grad[0]:
            dl2 = _f32c(dlin.reshape(-1, out_f))
            x2 = _f32c(x.reshape(-1, in_f)) if ctx.mask_in else None
            dx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            check(lib.feta_linear_dx(_ptr(dl2), _ptr(w), None, _ptr(x2), _ptr(dx), dl2.shape[0], in_f, out_f,
                                     _stream()), "feta_linear_dx")
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = _linear_wgrad(dlin, x, out_f, in_f, ctx.has_bias)
        return dx, dw, db, dz.view(z.shape), None, dg, dbeta, None, None


def linear_layernorm_impl(in_f, out_f):
    """0: separate projection + add_layer_norm launches; else the kernel family of the fused launch."""
    lib = _lib.load()
    if LINEAR_TC5 and lib.feta_linear_layernorm_supported(int(in_f), int(out_f)):
        return _IMPL_TC5
    if LINEAR_SIMT and LINEAR_LN_FUSED and lib.feta_linear_layernorm_simt_supported(int(in_f), int(out_f)):
        return _IMPL_SIMT
    return 0


def linear_layernorm_enabled(in_f, out_f):
    return linear_layernorm_impl(in_f, out_f) != 0


def linear_add_layer_norm(x, weight, bias, res, gamma, beta, eps=1e-5, bscale=None, mask_input_grad=False):
    return LinearAddLayerNormFn.apply(x, weight, bias, res, bscale, gamma, beta, eps, mask_input_grad)


class AddBatchNormFn(torch.autograd.Function):
    """y = BatchNorm1d(a + bscale * b) in TRAINING mode over rows (the layer's flattened [Nmax*B, d], padding rows
    included like the reference); ``roww`` (0/1 per row, optional) excludes rows from the statistics -- the
    static-shape layout pads beyond the batch maximum the reference pads to.  Running statistics are updated in
    place like ``nn.BatchNorm1d``.  Two launches each way (csrc/batchnorm.cu)."""

    @staticmethod
    def forward(ctx, a, b, bscale, roww, gamma, beta, running_mean, running_var, num_batches, momentum, eps):
        _need_cuda(a, b, gamma, beta)
        lib = _lib.load()
        a = _f32c(a)
        b = None if b is None else _f32c(b)
        bscale = None if bscale is None else _f32c(bscale)
        roww = None if roww is None else _f32c(roww)
        D = a.shape[-1]
        T = a.numel() // D
        y, z = torch.empty_like(a), torch.empty_like(a)
        mean = torch.empty(D, dtype=torch.float32, device=a.device)
        rstd = torch.empty(D, dtype=torch.float32, device=a.device)
        nblk = lib.feta_add_batchnorm_blocks(T)
        partial = torch.empty(nblk * 3 * D, dtype=torch.float32, device=a.device)
        gamma, beta = _f32c(gamma), _f32c(beta)
        check(lib.feta_add_batchnorm_fwd(_ptr(a), _ptr(b), _ptr(bscale), _ptr(roww), _ptr(gamma), _ptr(beta), _ptr(y),
                                         _ptr(z), _ptr(mean), _ptr(rstd), _ptr(running_mean), _ptr(running_var),
                                         _ptr(num_batches), _ptr(partial), float(momentum), float(eps), T, D, _stream()),
              "feta_add_batchnorm_fwd")
        ctx.save_for_backward(z, mean, rstd, gamma, bscale, roww)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        z, mean, rstd, gamma, bscale, roww = ctx.saved_tensors
        D = z.shape[-1]
        T = z.numel() // D
        dy = _f32c(dy)
        dz = torch.empty_like(z)
        dbs = torch.empty_like(z) if (ctx.has_b and bscale is not None) else None
        dg = torch.empty(D, dtype=torch.float32, device=z.device)
        db = torch.empty(D, dtype=torch.float32, device=z.device)
        nblk = lib.feta_add_batchnorm_blocks(T)
        partial = torch.empty(nblk * 3 * D, dtype=torch.float32, device=z.device)
        check(lib.feta_add_batchnorm_bwd(_ptr(dy), _ptr(z), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(bscale), _ptr(roww),
                                         _ptr(dz), _ptr(dbs), _ptr(dg), _ptr(db), _ptr(partial), T, D, _stream()),
              "feta_add_batchnorm_bwd")
        grad_b = None
        if ctx.has_b:
            grad_b = dbs if dbs is not None else dz
        return dz, grad_b, None, None, dg, db, None, None, None, None, None


def batchnorm_supported(D):
    return D % 4 == 0 and D <= 256 and 256 % D == 0


def add_batch_norm(a, b, bn, bscale=None, roww=None):
    """``bn``: an ``nn.BatchNorm1d`` in training mode (parameters, running buffers, momentum, eps are its own)."""
    momentum = 0.1 if bn.momentum is None else bn.momentum
    track = bn.track_running_stats and bn.running_mean is not None
    return AddBatchNormFn.apply(a, b, bscale, roww, bn.weight, bn.bias, bn.running_mean if track else None,
                                bn.running_var if track else None, bn.num_batches_tracked if track else None,
                                momentum, bn.eps)
