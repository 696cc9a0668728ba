"""torch.autograd bindings of the C-ABI kernels.

Every function here launches hand-written sm_100a kernels through ``_lib.call``; inputs must be
CUDA fp32 tensors.  There is no CPU or pure-PyTorch fallback: a CPU tensor raises.
"""
import os

import torch

from . import _lib
from .index import SegmentPlan

LEAKY_SLOPE = 0.2  # GATv2Conv default negative_slope (the reference never overrides it)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gasfm_b200 ops run on CUDA tensors only (no CPU fallback); got a tensor on %s" % t.device)
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("gasfm_b200 ops are fp32; got %s" % t.dtype)


def _rows(t):
    """(tensor, row stride) of a 2-D tensor whose rows are contiguous; copies otherwise."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    if t.shape[1] == 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t, t.stride(0) if t.shape[0] > 1 else t.shape[1]


# ---------------------------------------------------------------------------------------------
# fused GATv2 edge attention
# ---------------------------------------------------------------------------------------------
class _GatEdge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, XL, XR, att, bias, plan, heads, lazy_xl=None, rowmax_slot=None):
        _require_cuda(XL, XR, att, bias)
        XL, ldxl = _rows(XL)
        hc = XL.shape[1]
        head_dim = hc // heads
        T = plan.n_seg
        bcast = XR.shape[0] == 1 and T != 1
        XR = XR.contiguous()
        ldxr = 0 if bcast else hc
        att_flat = att.reshape(-1).contiguous()
        dev = XL.device
        out = torch.empty((T, hc), dtype=torch.float32, device=dev)
        seg_max = torch.empty((T, heads), dtype=torch.float32, device=dev)
        seg_sum = torch.empty((T, heads), dtype=torch.float32, device=dev)
        ws = None
        if plan.chunk > 0:
            ws = plan.workspace(_lib.size_query("gasfm_gat_ws_bytes", plan.max_chunks, heads, head_dim), dev)
        with _lib.device_guard(dev):
            _lib.call("gasfm_gat_edge_fwd", _lib.ptr(XL), ldxl, _lib.ptr(XR), ldxr, _lib.ptr(att_flat), _lib.ptr(bias),
                      *plan.abi_args(), heads, head_dim, LEAKY_SLOPE, 1,
                      _lib.ptr(out), _lib.ptr(seg_max), _lib.ptr(seg_sum), _lib.ptr(ws), _lib.stream_ptr())
        # lazy_xl: the projected sources are not kept for backward but rebuilt by the layer's EdgeBlockContext
        ctx.save_for_backward(XL if lazy_xl is None else None, XR, att_flat, bias, out, seg_max, seg_sum)
        ctx.plan, ctx.heads, ctx.bcast, ctx.att_shape, ctx.lazy_xl = plan, heads, bcast, att.shape, lazy_xl
        ctx.rowmax_slot = rowmax_slot
        return out

    @staticmethod
    def backward(ctx, d_out):
        XL, XR, att_flat, bias, out, seg_max, seg_sum = ctx.saved_tensors
        if XL is None:
            XL, _ = _rows(ctx.lazy_xl())
        plan, heads = ctx.plan, ctx.heads
        hc = XL.shape[1]
        head_dim = hc // heads
        dev = XL.device
        d_out = d_out.contiguous()
        out_nobias = out if bias is None else out - bias
        dXL, dXR, datt, rowmax = _gat_backward_launch(XL, XR, 0 if ctx.bcast else hc, att_flat, out_nobias, seg_max, seg_sum, d_out,
                                                      plan, heads, want_rowmax=ctx.rowmax_slot is not None)
        if ctx.rowmax_slot is not None:
            ctx.rowmax_slot[0].note_rowmax(ctx.rowmax_slot[1], dXL, rowmax)
        if ctx.bcast:
            dXR = dXR.sum(dim=0, keepdim=True)
        d_bias = None if bias is None else col_sum(d_out)
        return dXL, dXR, datt.view(ctx.att_shape), d_bias, None, None, None, None


def _gat_backward_launch(XL, XR, ldxr, att_flat, out_nobias, seg_max, seg_sum, d_out, plan, heads, want_rowmax=False):
    """-> (dXL, dXR, datt, rowmax | None).  ``want_rowmax``: also emit max |dXL[e, :]| per edge row (fp32, vectorised head shapes):
    the row scale of the fp16 input-gradient GEMM that reads dXL next."""
    hc = XL.shape[1]
    head_dim = hc // heads
    dev = XL.device
    bf16 = XL.dtype == torch.bfloat16
    covers_all = plan.perm is None or plan.n_edges == XL.shape[0]
    dXL = (torch.empty if covers_all else torch.zeros)((XL.shape[0], hc), dtype=XL.dtype, device=dev)
    dXR = torch.empty((plan.n_seg, hc), dtype=torch.float32, device=dev)
    datt = torch.empty(hc, dtype=torch.float32, device=dev)
    ws = plan.workspace(_lib.size_query("gasfm_gat_bwd_ws_bytes", XL.shape[0], plan.n_seg, plan.max_chunks, heads, head_dim), dev)
    rowmax = None
    if want_rowmax and not bf16 and _lib.load().gasfm_gat_edge_bwd_rowmax_supported(heads, head_dim):
        rowmax = (torch.empty if covers_all else torch.zeros)(XL.shape[0], dtype=torch.float32, device=dev)
    ldxl = XL.stride(0) if XL.shape[0] > 1 else hc
    args = (_lib.ptr(XL), ldxl, _lib.ptr(XR), ldxr, _lib.ptr(att_flat), _lib.ptr(out_nobias), _lib.ptr(seg_max), _lib.ptr(seg_sum),
            _lib.ptr(d_out), *plan.abi_args(), heads, head_dim, LEAKY_SLOPE, _lib.ptr(dXL), hc, _lib.ptr(dXR), _lib.ptr(datt))
    with _lib.device_guard(dev):
        if rowmax is not None:
            _lib.call("gasfm_gat_edge_bwd_rowmax", *args, _lib.ptr(rowmax), _lib.ptr(ws), _lib.stream_ptr())
        else:
            _lib.call("gasfm_gat_edge_bwd_bf16" if bf16 else "gasfm_gat_edge_bwd", *args, _lib.ptr(ws), _lib.stream_ptr())
    return dXL, dXR, datt, rowmax


def gat_edge_attention(XL, XR, att, bias, plan: SegmentPlan, heads: int, lazy_xl=None, rowmax_slot=None):
    """out[T,HC] = GATv2 softmax-aggregate of XL rows over ``plan``'s segments (+ bias).

    XL [E,HC] projected sources (rows may be a strided slice), XR [T,HC] projected targets, or
    [1,HC] to broadcast one query row to every target (stateless first block)."""
    return _GatEdge.apply(XL, XR, att, bias, plan, heads, lazy_xl, rowmax_slot)


def gat_edge_partial(XL, XR, att, plan: SegmentPlan, heads: int):
    """Un-normalised per-shard result (no autograd): (sum_e exp(s-max)*XL[e], max, sum).
    Used by the track-sharded multi-GPU path, which merges these across ranks."""
    bf16 = XL.dtype == torch.bfloat16          # bf16-STORED sources (kernel sweep, BASELINE.json configs[4]); arithmetic stays fp32
    _require_cuda(None if bf16 else XL, XR, att)
    XL, ldxl = _rows(XL)
    hc = XL.shape[1]
    head_dim = hc // heads
    T = plan.n_seg
    bcast = XR.shape[0] == 1 and T != 1
    XR = XR.contiguous()
    att_flat = att.reshape(-1).contiguous()
    dev = XL.device
    out = torch.empty((T, hc), dtype=torch.float32, device=dev)
    seg_max = torch.empty((T, heads), dtype=torch.float32, device=dev)
    seg_sum = torch.empty((T, heads), dtype=torch.float32, device=dev)
    ws = None
    if plan.chunk > 0:
        ws = plan.workspace(_lib.size_query("gasfm_gat_ws_bytes", plan.max_chunks, heads, head_dim), dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_gat_edge_fwd_bf16" if bf16 else "gasfm_gat_edge_fwd", _lib.ptr(XL), ldxl, _lib.ptr(XR), 0 if bcast else hc, _lib.ptr(att_flat), None,
                  *plan.abi_args(), heads, head_dim, LEAKY_SLOPE, 0,
                  _lib.ptr(out), _lib.ptr(seg_max), _lib.ptr(seg_sum), _lib.ptr(ws), _lib.stream_ptr())
    return out, seg_max, seg_sum


def gat_edge_backward_raw(XL, XR, att, out_nobias, seg_max, seg_sum, d_out, plan, heads, want_rowmax=False):
    """Backward kernel with explicitly supplied (global) softmax statistics; returns (dXL, dXR, datt) -- plus the row maxima of
    dXL with ``want_rowmax``.  The multi-GPU path calls this with the merged statistics.  bf16 ``XL`` -> bf16 ``dXL``."""
    XL, _ = _rows(XL)
    hc = XL.shape[1]
    bcast = XR.shape[0] == 1 and plan.n_seg != 1
    dXL, dXR, datt, rowmax = _gat_backward_launch(XL, XR.contiguous(), 0 if bcast else hc, att.reshape(-1).contiguous(),
                                                  out_nobias.contiguous(), seg_max.contiguous(), seg_sum.contiguous(),
                                                  d_out.contiguous(), plan, heads, want_rowmax)
    if bcast:
        dXR = dXR.sum(dim=0, keepdim=True)
    return (dXL, dXR, datt, rowmax) if want_rowmax else (dXL, dXR, datt)


# ---------------------------------------------------------------------------------------------
# LayerNorm + ReLU on observation features
# ---------------------------------------------------------------------------------------------
COL_SUM_MIN_ROWS = 1024   # below this torch's own reduction is as fast


def col_sum(x, keepdim=False):
    """``x.sum(dim=0)`` for a [rows, width] fp32 CUDA matrix (bias gradients) -- deterministic two-stage kernel."""
    if not (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.shape[0] >= COL_SUM_MIN_ROWS and x.stride(1) == 1):
        return x.sum(dim=0, keepdim=keepdim)
    rows, w = x.shape
    out = torch.empty(w, dtype=torch.float32, device=x.device)
    nbytes = _lib.size_query("gasfm_col_sum_ws_bytes", rows, w)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device) if nbytes else None
    with _lib.device_guard(x.device):
        _lib.call("gasfm_col_sum", _lib.ptr(x), x.stride(0), rows, w, _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr())
    return out.unsqueeze(0) if keepdim else out


def ln_relu_width_supported(width):
    """Mirror of the kernel's own predicate (gasfm_ln_relu_fwd / _bwd, csrc/edge_ops.cu): rows up to 1024 wide as
    float4, or up to 256 wide scalar."""
    return 0 < width <= 1024 and (width % 4 == 0 or width <= 256)


def _ln_relu_forward(x, gamma, beta, eps):
    _require_cuda(x, gamma, beta)
    x = x.contiguous()
    E, w = x.shape
    dev = x.device
    y = torch.empty_like(x)
    mean = rstd = None
    if gamma is not None:
        gamma, beta = gamma.contiguous(), beta.contiguous()
        mean = torch.empty(E, dtype=torch.float32, device=dev)
        rstd = torch.empty(E, dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_ln_relu_fwd", _lib.ptr(x), E, w, _lib.ptr(gamma), _lib.ptr(beta), float(eps),
                  _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd), _lib.stream_ptr())
    return x, y, mean, rstd, gamma, beta


def _ln_relu_backward(x, mean, rstd, gamma, beta, dy, add=None):
    E, w = x.shape
    dev = x.device
    dy = dy.contiguous()
    dx = torch.empty_like(x)
    dgamma = dbeta = ws = None
    if gamma is not None:
        dgamma = torch.empty(w, dtype=torch.float32, device=dev)
        dbeta = torch.empty(w, dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, _lib.size_query("gasfm_ln_relu_bwd_ws_bytes", E, w) // 4), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_ln_relu_bwd", _lib.ptr(dy), _lib.ptr(x), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
                  _lib.ptr(beta), _lib.ptr(None if add is None else add.contiguous()), E, w, _lib.ptr(dx),
                  _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(ws), _lib.stream_ptr())
    return dx, dgamma, dbeta


class _LnRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x, y, mean, rstd, gamma, beta = _ln_relu_forward(x, gamma, beta, eps)
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        dx, dgamma, dbeta = _ln_relu_backward(x, mean, rstd, gamma, beta, dy)
        return dx, dgamma, dbeta, None


class _LnReluSkip(torch.autograd.Function):
    """(relu(LN(x)), x) as ONE autograd node: the second output is x itself, to be consumed by the residual
    branch of the layer.  Backward gets both gradients at once and the LN+ReLU backward kernel adds the skip
    gradient while writing dx -- no separate [E,d] gradient-accumulation pass."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x, y, mean, rstd, gamma, beta = _ln_relu_forward(x, gamma, beta, eps)
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        ctx.set_materialize_grads(False)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        if dy is None:
            zeros = None if gamma is None else torch.zeros_like(gamma)
            return dskip, zeros, zeros, None
        dx, dgamma, dbeta = _ln_relu_backward(x, mean, rstd, gamma, beta, dy, add=dskip)
        return dx, dgamma, dbeta, None


def ln_relu_with_skip(x, gamma=None, beta=None, eps=1e-5):
    """-> (relu(layer_norm(x)), x): use the second value wherever the un-normalised input is needed again."""
    return _LnReluSkip.apply(x, gamma, beta, eps)


def ln_relu(x, gamma=None, beta=None, eps=1e-5):
    """relu(layer_norm(x)) over the last dim of an [E,d] matrix; gamma=None -> relu(x)."""
    return _LnRelu.apply(x, gamma, beta, eps)


# ---------------------------------------------------------------------------------------------
# row / column pooling
# ---------------------------------------------------------------------------------------------
def seg_sum_raw(X, plan: SegmentPlan, scale=1.0, mean=False):
    _require_cuda(X)
    X, ldx = _rows(X)
    w = X.shape[1]
    dev = X.device
    out = torch.empty((plan.n_seg, w), dtype=torch.float32, device=dev)
    ws = None
    if plan.chunk > 0:
        ws = plan.workspace(_lib.size_query("gasfm_seg_sum_ws_bytes", plan.max_chunks, w), dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_seg_sum", _lib.ptr(X), ldx, w, *plan.abi_args(), float(scale), int(mean),
                  _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr())
    return out


class _SegPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, plan, seg_of_edge, scale, mean):
        ctx.plan, ctx.seg_of_edge, ctx.scale, ctx.mean, ctx.n_rows = plan, seg_of_edge, scale, mean, X.shape[0]
        return seg_sum_raw(X, plan, scale, mean)

    @staticmethod
    def backward(ctx, d_out):
        d_out = d_out.contiguous()
        w = d_out.shape[1]
        dev = d_out.device
        dX = torch.empty((ctx.n_rows, w), dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            _lib.call("gasfm_seg_bcast", _lib.ptr(d_out), w, _lib.ptr(ctx.seg_of_edge), _lib.ptr(ctx.plan.seg_ptr),
                      ctx.n_rows, float(ctx.scale), int(ctx.mean), _lib.ptr(dX), _lib.stream_ptr())
        return dX, None, None, None, None


def seg_pool(X, plan, seg_of_edge, scale=1.0, mean=False):
    """Differentiable segment sum / mean of the rows of X [E,w] -> [T,w]."""
    return _SegPool.apply(X, plan, seg_of_edge, scale, mean)


# ---------------------------------------------------------------------------------------------
# per-observation update: out = pscale*P + scale*(x0 @ W0^T + S[col] + V[row] + g) + skip
# ---------------------------------------------------------------------------------------------
class _EdgeUpdate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, P, x0, W0, S, V, g, skip, index, pscale, scale, rowmax_slot=None):
        _require_cuda(P, x0, W0, S, V, g, skip)
        P, ldp = _rows(P)
        E, w = P.shape
        dev = P.device
        d0 = 0
        if x0 is not None:
            x0, W0 = x0.contiguous(), W0.contiguous()
            d0 = x0.shape[1]
        if skip is not None:
            skip, ldskip = _rows(skip)
        else:
            ldskip = w
        S = None if S is None else S.contiguous()
        V = None if V is None else V.contiguous()
        g = None if g is None else g.contiguous()
        out = torch.empty((E, w), dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            _lib.call("gasfm_edge_update_fwd", _lib.ptr(P), ldp, _lib.ptr(x0), d0, _lib.ptr(W0), _lib.ptr(S), _lib.ptr(V),
                      _lib.ptr(g), _lib.ptr(skip), ldskip, _lib.ptr(index.row_idx), _lib.ptr(index.col_idx), E, w,
                      float(pscale), float(scale), _lib.ptr(out), _lib.stream_ptr())
        ctx.save_for_backward(x0, W0)
        ctx.index, ctx.pscale, ctx.scale = index, pscale, scale
        ctx.rowmax_slot = rowmax_slot
        ctx.has = (S is not None, V is not None, g is not None, skip is not None)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x0, W0 = ctx.saved_tensors
        index, scale = ctx.index, ctx.scale
        has_S, has_V, has_g, has_skip = ctx.has
        d_out = d_out.contiguous()
        dP = d_out if ctx.pscale == 1.0 else d_out * ctx.pscale
        dS = seg_sum_raw(d_out, index.by_track, scale) if (has_S and ctx.needs_input_grad[3]) else None
        E, w = d_out.shape
        need_dV = (has_V and ctx.needs_input_grad[4]) or (has_g and ctx.needs_input_grad[5])
        need_x0 = x0 is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        x0_kernel = need_x0 and w % 4 == 0 and w <= 1024 and 1 <= x0.shape[1] <= 4
        slot = ctx.rowmax_slot if ctx.pscale == 1.0 else None            # dP is d_out itself only then
        dV = dx0 = dW0 = None
        if need_dV and x0_kernel and UPDATE_BWD_FUSED and index.by_view.chunk > 0:
            # ONE storage-order pass over d_out: per-view sums, the rank-d0 term's gradients and the row maxima
            dV, dx0, dW0, rowmax = _update_backward_views(d_out, x0, W0, scale, index.by_view, want_rowmax=slot is not None)
            if slot is not None:
                slot[0].note_rowmax(slot[1], d_out, rowmax)
        else:
            if need_dV:
                dV = seg_sum_raw(d_out, index.by_view, scale)
            if x0_kernel:
                dx0, dW0, rowmax = _x0_backward(d_out, x0, W0, scale, want_rowmax=slot is not None)
                if slot is not None:
                    slot[0].note_rowmax(slot[1], d_out, rowmax)
            elif need_x0:
                dx0 = torch.mm(d_out, W0).mul_(scale)
                dW0 = torch.mm(d_out.t(), x0).mul_(scale)
        if dV is not None:
            shard = getattr(index, "shard", None)
            if shard is not None and shard.world > 1:
                # track-sharded scene: V and g are replicated, this rank saw only its own observations of every view
                dV = shard.exchange.allreduce_sum(dV)
        dg = col_sum(dV, keepdim=True) if (has_g and ctx.needs_input_grad[5]) else None
        return (dP, dx0, dW0, dS, dV if has_V else None, dg, d_out if has_skip else None, None, None, None, None)


def edge_update(P, x0, W0, S, V, g, skip, index, pscale=1.0, scale=0.25, rowmax_slot=None):
    return _EdgeUpdate.apply(P, x0, W0, S, V, g, skip, index, pscale, scale, rowmax_slot)


# ---------------------------------------------------------------------------------------------
# dense per-observation projection on the tensor cores (tcgen05, 3xTF32 split)
# ---------------------------------------------------------------------------------------------
def _split_tf32(w):
    w = w.contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    with _lib.device_guard(w.device):
        _lib.call("gasfm_split_tf32", _lib.ptr(w), _lib.ptr(hi), _lib.ptr(lo), w.numel(), _lib.stream_ptr())
    return hi, lo


def gemm_tf32x3_supported(M, N, K, lda, ldc):
    return bool(_lib.load().gasfm_linear_tf32x3_supported(int(M), int(N), int(K), int(lda), int(ldc)))


def _split_f16(w):
    """Rows of w scaled by a power of two and split into fp16 hi + lo (see gemm_f16x2.cu) -> (hi, lo, descale[N])."""
    w = w.contiguous()
    hi = torch.empty(w.shape, dtype=torch.float16, device=w.device)
    lo = torch.empty(w.shape, dtype=torch.float16, device=w.device)
    descale = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
    with _lib.device_guard(w.device):
        _lib.call("gasfm_split_f16", _lib.ptr(w), w.shape[0], w.shape[1], _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(descale),
                  _lib.stream_ptr())
    return hi, lo, descale


def gemm_f16x2_supported(M, N, K, lda, ldc):
    return bool(_lib.load().gasfm_linear_f16x2_supported(int(M), int(N), int(K), int(lda), int(ldc)))


# "f16x2": scaled 2 x FP16 split where the shape allows (K <= 256), 3xTF32 otherwise;  "tf32x3": always 3xTF32
GEMM_KIND = os.environ.get("GASFM_GEMM", "f16x2")


def gemm_tc(a, b, bias=None, out=None, accumulate=False, kind=None):
    """a [M,K] (rows contiguous, any row stride) times b[N,K]^T (+ bias[N]) -> [M,N] on the tcgen05 tensor cores
    with fp32-level accuracy (split-operand products).  ``out`` + ``accumulate``: out += a b^T (used to sum the
    input gradients of projections sharing x)."""
    a, lda = _rows(a)
    M, K = a.shape
    N = b.shape[0]
    kind = kind or GEMM_KIND
    c = torch.empty((M, N), dtype=torch.float32, device=a.device) if out is None else out
    bias_ptr = _lib.ptr(None if bias is None else bias.contiguous())
    if kind == "f16x2" and gemm_f16x2_supported(M, N, K, lda, N):
        hi, lo, descale = _split_f16(b)
        with _lib.device_guard(a.device):
            _lib.call("gasfm_linear_f16x2", _lib.ptr(a), lda, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(descale), bias_ptr,
                      _lib.ptr(c), N, M, N, K, 1, int(bool(accumulate)), None, _lib.stream_ptr())
        return c
    hi, lo = _split_tf32(b)
    with _lib.device_guard(a.device):
        _lib.call("gasfm_linear_tf32x3", _lib.ptr(a), lda, _lib.ptr(hi), _lib.ptr(lo), bias_ptr, _lib.ptr(c), N, M, N, K,
                  int(bool(accumulate)), _lib.stream_ptr())
    return c


F16X2_MAX_GROUPS = 3


def gemm_f16x2_groups(a, weights, biases, want_amax=False):
    """Several projections y_g = a W_g^T + b_g of the SAME a (equal output widths) in one kernel: a is read from
    HBM once.  Returns the [M, G*N] buffer; y_g = out[:, g*N:(g+1)*N].  ``want_amax``: also return max|a| as a
    1-element device tensor (the producers see every row maximum anyway; it scales the fp16 weight gradient)."""
    a, lda = _rows(a)
    M, K = a.shape
    G, N = len(weights), weights[0].shape[0]
    hi, lo, descale = _split_f16(torch.cat(list(weights), dim=0))
    bias = torch.cat([b if b is not None else torch.zeros(N, dtype=torch.float32, device=a.device) for b in biases])
    c = torch.empty((M, G * N), dtype=torch.float32, device=a.device)
    amax = torch.empty(1, dtype=torch.float32, device=a.device) if want_amax else None
    with _lib.device_guard(a.device):
        _lib.call("gasfm_linear_f16x2", _lib.ptr(a), lda, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(descale), _lib.ptr(bias),
                  _lib.ptr(c), G * N, M, N, K, G, 0, _lib.ptr(amax), _lib.stream_ptr())
    return (c, amax) if want_amax else c


def gemm_f16x2_ln_y_supported(M, N, K, lda, ldc):
    return bool(_lib.load().gasfm_linear_f16x2_ln_y_supported(int(M), int(N), int(K), int(lda), int(ldc)))


def gemm_f16x2_groups_ln(x_raw, gamma, beta, eps, weights, biases, want_y=False):
    """gemm_f16x2_groups over relu(layer_norm(x_raw)) with the normalisation done inside the GEMM's operand producer:
    x_raw is read once.  -> (out [M, G*N], max|operand| [1], mean [M], rstd [M]) and, with ``want_y`` (N = K = 256 kernel
    only, ``gemm_f16x2_ln_y_supported``), the normalised matrix itself as a by-product of the same pass."""
    x_raw, lda = _rows(x_raw)
    M, K = x_raw.shape
    G, N = len(weights), weights[0].shape[0]
    hi, lo, descale = _split_f16(torch.cat(list(weights), dim=0))
    bias = torch.cat([b if b is not None else torch.zeros(N, dtype=torch.float32, device=x_raw.device) for b in biases])
    dev = x_raw.device
    c = torch.empty((M, G * N), dtype=torch.float32, device=dev)
    amax = torch.empty(1, dtype=torch.float32, device=dev)
    mean = torch.empty(M, dtype=torch.float32, device=dev)
    rstd = torch.empty(M, dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        if want_y:
            y = torch.empty((M, K), dtype=torch.float32, device=dev)
            _lib.call("gasfm_linear_f16x2_ln_y", _lib.ptr(x_raw), lda, _lib.ptr(gamma.contiguous()), _lib.ptr(beta.contiguous()),
                      float(eps), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(y), K, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(descale),
                      _lib.ptr(bias), _lib.ptr(c), G * N, M, N, K, G, _lib.ptr(amax), _lib.stream_ptr())
            return c, amax, mean, rstd, y
        _lib.call("gasfm_linear_f16x2_ln", _lib.ptr(x_raw), lda, _lib.ptr(gamma.contiguous()), _lib.ptr(beta.contiguous()), float(eps),
                  _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(descale), _lib.ptr(bias), _lib.ptr(c), G * N,
                  M, N, K, G, _lib.ptr(amax), _lib.stream_ptr())
    return c, amax, mean, rstd


def gemm_tf32x3(a, b, bias=None, out=None, accumulate=False):
    return gemm_tc(a, b, bias, out, accumulate, kind="tf32x3")


def gemm_tf32x3_cat(a_list, b, bias=None, want_amax=False):
    """[A_0 | A_1 | ..] b^T for 1..4 matrices A_i [M, seg_k] that live in separate buffers; b is [N, n * seg_k].
    One pass over every A_i and one write of the result (3xTF32: no per-row scale to agree on across the A_i).
    ``want_amax``: also return max|A_i| per segment as an n-element device tensor."""
    import ctypes

    rows = [_rows(a) for a in a_list]
    M, seg_k = rows[0][0].shape
    N = b.shape[0]
    n = len(rows)
    assert b.shape[1] == n * seg_k and all(r[0].shape == (M, seg_k) for r in rows)
    hi, lo = _split_tf32(b)
    c = torch.empty((M, N), dtype=torch.float32, device=b.device)
    ptrs = (ctypes.c_void_p * n)(*[r[0].data_ptr() for r in rows])
    lds = (ctypes.c_int64 * n)(*[r[1] for r in rows])
    amax = torch.empty(n, dtype=torch.float32, device=b.device) if want_amax else None
    with _lib.device_guard(b.device):
        _lib.call("gasfm_linear_tf32x3_cat", ptrs, lds, n, seg_k, _lib.ptr(hi), _lib.ptr(lo),
                  _lib.ptr(None if bias is None else bias.contiguous()), _lib.ptr(c), N, M, N, 0, _lib.ptr(amax),
                  _lib.stream_ptr())
    return (c, amax) if want_amax else c


# concatenated input gradient of a block: "f16x2" (default) = the fp16 GEMM in its ONE-pass form whenever the kernels that
# produced the three output gradients left their row maxima behind (EdgeBlockContext.note_rowmax), 3xTF32 otherwise;
# "tf32x3" = always 3xTF32.  (The two-pass fp16 form without row maxima is slower than 3xTF32, profiles/r02_fusion_ab.md.)
DX_KIND = os.environ.get("GASFM_DX", "f16x2")


def gemm_f16x2_cat_supported(M, N, n_seg, seg_k):
    return bool(_lib.load().gasfm_linear_f16x2_cat_supported(int(M), int(N), int(n_seg), int(seg_k), int(N)))


def gemm_cat(a_list, b, bias=None, want_amax=False):
    """[A_0 | A_1 | ..] b^T on the tensor cores when no row maxima of the A_i are at hand: 3xTF32."""
    return gemm_tf32x3_cat(a_list, b, bias, want_amax)


def gemm_f16x2_cat(a_list, b, bias=None, want_amax=False, rowmax=None):
    """gemm_tf32x3_cat on the scaled 2 x FP16 path (twice the tensor rate; one row scale across the segments, found in a first
    pass over the tile that L2 absorbs).  Same contract: returns c [M,N] (and max|A_i| per segment)."""
    import ctypes

    rows = [_rows(a) for a in a_list]
    M, seg_k = rows[0][0].shape
    N, n = b.shape[0], len(rows)
    assert b.shape[1] == n * seg_k and all(r[0].shape == (M, seg_k) for r in rows)
    hi, lo, descale = _split_f16(b)
    c = torch.empty((M, N), dtype=torch.float32, device=b.device)
    ptrs = (ctypes.c_void_p * n)(*[r[0].data_ptr() for r in rows])
    lds = (ctypes.c_int64 * n)(*[r[1] for r in rows])
    amax = torch.empty(n, dtype=torch.float32, device=b.device) if want_amax else None
    bias_ptr = _lib.ptr(None if bias is None else bias.contiguous())
    with _lib.device_guard(b.device):
        if rowmax is not None:
            # ``rowmax``: per-segment [M] row maxima left behind by the kernels that produced the segments -> single-pass producer
            rms = (ctypes.c_void_p * n)(*[r.data_ptr() for r in rowmax])
            _lib.call("gasfm_linear_f16x2_cat_rowmax", ptrs, lds, rms, n, seg_k, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(descale),
                      bias_ptr, _lib.ptr(c), N, M, N, _lib.ptr(amax), _lib.stream_ptr())
        else:
            _lib.call("gasfm_linear_f16x2_cat", ptrs, lds, n, seg_k, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(descale),
                      bias_ptr, _lib.ptr(c), N, M, N, _lib.ptr(amax), _lib.stream_ptr())
    return (c, amax) if want_amax else c


def gemm_f16x2(a, b, bias=None, out=None, accumulate=False):
    return gemm_tc(a, b, bias, out, accumulate, kind="f16x2")


def wgrad_tf32x3_supported(E, n_out, k_out, lddy, ldx):
    lib = _lib.load()
    return bool(lib.gasfm_wgrad_small_supported(int(n_out), int(k_out), int(lddy), int(ldx)) or
                lib.gasfm_wgrad_tf32x3_supported(int(E), int(n_out), int(k_out), int(lddy), int(ldx)))


def wgrad_tf32x3(dy, x, with_bias=False):
    """dW[Nout,Kout] = dy[E,Nout]^T @ x[E,Kout]: tensor cores (3xTF32, deterministic split-K), or the SIMT
    register-tiled kernel for the narrow shipped widths (32 / 64).  ``with_bias`` also returns the column
    sums of dy (the bias gradient), accumulated in the same pass."""
    dy, lddy = _rows(dy)
    x, ldx = _rows(x)
    E, n_out = dy.shape
    k_out = x.shape[1]
    dev = dy.device
    dw = torch.empty((n_out, k_out), dtype=torch.float32, device=dev)
    db = torch.empty(n_out, dtype=torch.float32, device=dev) if with_bias else None
    small = _lib.load().gasfm_wgrad_small_supported(n_out, k_out, lddy, ldx)
    name = "gasfm_wgrad_small" if small else "gasfm_wgrad_tf32x3"
    ws = torch.empty(_lib.size_query(name + "_ws_bytes", n_out, k_out) // 4, dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call(name, _lib.ptr(dy), lddy, _lib.ptr(x), ldx, E, n_out, k_out, _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws),
                  _lib.stream_ptr())
    return (dw, db) if with_bias else dw


# "f16x2": fp16 weight gradient where both operand maxima are at hand (grouped projections), 3xTF32 otherwise
WGRAD_KIND = os.environ.get("GASFM_WGRAD", "f16x2")


def wgrad_f16x2_supported(E, n_out, k_out, lddy, ldx):
    return bool(_lib.load().gasfm_wgrad_f16x2_supported(int(E), int(n_out), int(k_out), int(lddy), int(ldx)))


def wgrad_f16x2(dy, x, amax_dy, amax_x, with_bias=False):
    """dW = dy^T x (and db) on the fp16 tensor-core path; amax_dy / amax_x: 1-element DEVICE tensors holding max|dy|,
    max|x| (any upper bound within a few binades works), as returned by gemm_tf32x3_cat / gemm_f16x2_groups."""
    dy, lddy = _rows(dy)
    x, ldx = _rows(x)
    E, n_out = dy.shape
    k_out = x.shape[1]
    dev = dy.device
    dw = torch.empty((n_out, k_out), dtype=torch.float32, device=dev)
    db = torch.empty(n_out, dtype=torch.float32, device=dev) if with_bias else None
    ws = torch.empty(_lib.size_query("gasfm_wgrad_f16x2_ws_bytes", n_out, k_out) // 4, dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_wgrad_f16x2", _lib.ptr(dy), lddy, _lib.ptr(x), ldx, _lib.ptr(amax_dy), _lib.ptr(amax_x), E, n_out, k_out,
                  _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), _lib.stream_ptr())
    return (dw, db) if with_bias else dw


def wgrad_f16x2_multi(dys, x, amax_dys, amax_x):
    """[dW_g, db_g] for 2..3 projections of the SAME x in one launch (x is read from HBM once, by the other groups from L2).
    ``amax_dys``: n-element device tensor (max |dY_g|), ``amax_x``: 1-element device tensor.  -> (dW [n,Nout,Kout], db [n,Nout])."""
    import ctypes

    rows = [_rows(dy) for dy in dys]
    x, ldx = _rows(x)
    n = len(rows)
    E, n_out = rows[0][0].shape
    k_out = x.shape[1]
    dev = x.device
    dw = torch.empty((n, n_out, k_out), dtype=torch.float32, device=dev)
    db = torch.empty((n, n_out), dtype=torch.float32, device=dev)
    ws = torch.empty(_lib.size_query("gasfm_wgrad_f16x2_ws_bytes", n_out, k_out) // 4, dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * n)(*[r[0].data_ptr() for r in rows])
    lds = (ctypes.c_int64 * n)(*[r[1] for r in rows])
    with _lib.device_guard(dev):
        _lib.call("gasfm_wgrad_f16x2_multi", ptrs, lds, n, _lib.ptr(x), ldx, _lib.ptr(amax_dys.contiguous()), _lib.ptr(amax_x),
                  E, n_out, k_out, _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), _lib.stream_ptr())
    return dw, db


WGRAD_MULTI = os.environ.get("GASFM_WGRAD_MULTI", "1") != "0"     # A/B switch: one launch per block instead of three


def _linear_backward(x, weight, dy, need_x, need_w, need_b, dx_out=None, amax=None):
    """Gradients of y = x W^T + b on the tensor-core kernels; ``dx_out``: accumulate dX into this buffer."""
    dy = dy.contiguous()
    dx = dw = db = None
    M, N = dy.shape
    K = weight.shape[1]
    if need_x:
        if gemm_tf32x3_supported(M, K, N, N, K):
            dx = gemm_tc(dy, weight.t(), out=dx_out, accumulate=dx_out is not None)   # dX = dY (W^T)^T
        elif dx_out is not None:
            dx = dx_out.add_(dy @ weight)
        else:
            dx = dy @ weight
    if need_w or need_b:
        ldx = x.stride(0) if x.stride(1) == 1 else K
        if amax is not None and WGRAD_KIND == "f16x2" and wgrad_f16x2_supported(M, N, K, N, ldx):
            dw, db = wgrad_f16x2(dy, x, amax[0], amax[1], with_bias=True)
        elif wgrad_tf32x3_supported(M, N, K, N, ldx):
            dw, db = wgrad_tf32x3(dy, x, with_bias=True)
        else:
            dw = dy.t() @ x
            db = col_sum(dy)
    return dx, dw, db


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return gemm_tc(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx, dw, db = _linear_backward(x, weight, dy, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                      ctx.has_bias and ctx.needs_input_grad[2])
        return dx, dw, (db if ctx.has_bias else None)


class _LinearMulti(torch.autograd.Function):
    """Several projections of the SAME input: y_i = x W_i^T + b_i.  Backward sums the input gradients
    inside the GEMM epilogue (dX += dY_i W_i) instead of materialising one dX per projection and adding."""

    @staticmethod
    def forward(ctx, x, *wb):
        weights, biases = wb[0::2], wb[1::2]
        ctx.save_for_backward(x, *weights)
        ctx.n = len(weights)
        M, K = x.shape
        N = weights[0].shape[0]
        lda = x.stride(0) if x.stride(1) == 1 else K
        if (GEMM_KIND == "f16x2" and 2 <= len(weights) <= F16X2_MAX_GROUPS and all(w.shape[0] == N for w in weights)
                and gemm_f16x2_supported(M, N, K, lda, len(weights) * N)):
            out, ctx.x_amax = gemm_f16x2_groups(x, weights, biases, want_amax=True)   # x is read once for all projections
            return tuple(out[:, g * N:(g + 1) * N] for g in range(len(weights)))
        ctx.x_amax = None
        return tuple(gemm_tc(x, w, b) for w, b in zip(weights, biases))

    @staticmethod
    def backward(ctx, *dys):
        x, *weights = ctx.saved_tensors
        grads = []
        dx = None
        need_x = ctx.needs_input_grad[0]
        M, K = x.shape
        n_out = weights[0].shape[0]
        fused_dx = (need_x and 2 <= len(weights) <= 4 and all(dy is not None for dy in dys)
                    and all(w.shape[0] == n_out for w in weights) and n_out % 4 == 0
                    and gemm_tf32x3_supported(M, K, len(weights) * n_out, n_out, K))
        if fused_dx:
            # dX = [dY_0 | dY_1 | ..] [W_0; W_1; ..]: one pass over every dY_i, one write of dX
            dys = [dy.contiguous() for dy in dys]
            dx, dy_amax = gemm_cat(dys, torch.cat([w.t() for w in weights], dim=1), want_amax=True)
        ldx = x.stride(0) if x.stride(1) == 1 else K
        if (fused_dx and ctx.x_amax is not None and WGRAD_MULTI and WGRAD_KIND == "f16x2" and len(weights) <= 3
                and all(ctx.needs_input_grad[1 + 2 * i] for i in range(len(weights)))
                and wgrad_f16x2_supported(M, n_out, K, n_out, ldx)):
            # every weight gradient of the node in one launch: x is read from HBM once
            dw_all, db_all = wgrad_f16x2_multi(dys, x, dy_amax, ctx.x_amax)
            for i in range(len(weights)):
                grads += [dw_all[i], db_all[i] if ctx.needs_input_grad[2 + 2 * i] else None]
            return (dx, *grads)
        for i, (w, dy) in enumerate(zip(weights, dys)):
            # both operand maxima are known (x from the forward GEMM, dY_i from the GEMM above): fp16 weight gradient
            amax = (dy_amax[i:i + 1], ctx.x_amax) if (fused_dx and ctx.x_amax is not None) else None
            g, dw, db = _linear_backward(x, w, dy, need_x and not fused_dx, ctx.needs_input_grad[1 + 2 * i],
                                         ctx.needs_input_grad[2 + 2 * i], dx_out=None if fused_dx else dx, amax=amax)
            if need_x and not fused_dx:
                dx = g
            grads += [dw, db]
        return (dx, *grads)


# ---------------------------------------------------------------------------------------------
# one GASFM block's observation-level front end: relu(LN(x_raw)) -> lin_l (x2) + lin_proj, as ONE autograd node
# ---------------------------------------------------------------------------------------------
# LayerNorm + ReLU inside the projection GEMM's operand producer (default: wherever the CTA-pair kernel runs, N = K = 256).
# That kernel converts the operand tile once for all three projections, so its producers have time to spare: the
# normalisation costs 5 % of the GEMM and the separate LayerNorm pass over [E, d] disappears; relu(LN(x)) is written out from
# the same registers only when backward needs it.  cfg3 at d = 256 (recompute): 570.9 -> 552 ms per step, and the step then
# fits a second graph capture (end-to-end 579.9 eager -> 553.5 ms replayed); cfg2: no change (profiles/r02_fusion_ab.md).
# "0": never; "force": also on the cta_group::1 kernel (other widths), where backward has to rebuild relu(LN(x)) separately.
_LN_FUSED_MODE = os.environ.get("GASFM_LN_FUSED", "1")
LN_FUSED = _LN_FUSED_MODE != "0"
LN_FUSED_ANY_SHAPE = _LN_FUSED_MODE == "force"
ACTIVATION_RECOMPUTE = os.environ.get("GASFM_RECOMPUTE", "auto")   # "on" | "off" | "auto" (decided per scene by the model)
# "auto" keeps the activations of as many of the LAST blocks as fit (their backward runs first and frees them before the
# recomputed blocks need room); GASFM_RECOMPUTE_KEEP=<k> fixes that number
RECOMPUTE_KEEP = os.environ.get("GASFM_RECOMPUTE_KEEP", "auto")
_recompute_now = False
_recompute_any = False
last_recompute_plan = (0, 0)      # (blocks recomputed, blocks) of the most recent forward with a policy decision


def recompute_plan(n_obs, width, n_blocks, total_memory, grad_enabled=True, mode=None, keep=None):
    """Activation recompute policy -> number of (leading) blocks that recompute (pure arithmetic, no device access).

    A block keeps five [E, width] fp32 tensors for backward (x_raw, the three grouped projections and relu(LN(x_raw))); with
    recompute only x_raw stays.  "auto": nothing is recomputed while the kept activations of all blocks take less than half of
    the device memory; beyond that the LAST blocks keep theirs as far as a conservative budget allows (their backward runs
    first and frees them before the recomputed blocks need room): 80 % of the device minus what a fully recomputed step was
    measured to hold at its peak -- 1.5 x_raw-sized tensors per block + 4 transient ones (cfg3 at d = 256: 101 GiB = 21.3 such
    tensors for 12 blocks) -- at four tensors per kept block.  "on": every block (unless ``keep`` says otherwise); "off": none."""
    mode = ACTIVATION_RECOMPUTE if mode is None else mode
    keep = RECOMPUTE_KEEP if keep is None else keep
    if mode == "off" or (mode == "auto" and not grad_enabled):
        return 0
    unit = float(n_obs) * width * 4
    if mode == "auto" and 5.0 * unit * n_blocks <= 0.5 * total_memory:
        return 0
    if keep != "auto":
        kept = int(keep)
    elif mode == "on":
        kept = 0
    else:
        kept = int((0.8 * total_memory - unit * (1.5 * n_blocks + 4.0)) // (4.0 * unit))
    return n_blocks - max(0, min(n_blocks, kept))


def set_activation_recompute(flag, first_of_forward=False):
    """Blocks whose forward is issued while this is set keep only ``x_raw`` + LayerNorm statistics and rebuild
    relu(LN(x_raw)) and the projected attention sources in backward (activation memory)."""
    global _recompute_now, _recompute_any
    _recompute_now = bool(flag)
    _recompute_any = bool(flag) if first_of_forward else (_recompute_any or bool(flag))


def activation_recompute_enabled():
    """Whether the block being issued recomputes its activations."""
    return _recompute_now


def activation_recompute_used():
    """Whether any block of the most recent forward pass recomputes its activations."""
    return _recompute_any


class EdgeBlockContext:
    """What one block needs to rebuild its [E, d] activations in backward: shared by the block's projection node and
    its two attention nodes (which fetch their ``XL`` through ``xl_getter``)."""

    def __init__(self, recompute):
        self.recompute = bool(recompute)
        self.ln_fused = False
        self.y = self.xl = None
        self.args = None
        self.rowmax = {}

    def slot(self, i):
        """Handle for the backward node that produces the i-th output gradient of the block's projections (0 / 1: the two
        attention kernels' dXL, 2: the observation update's dP): it leaves the row maxima of that gradient here."""
        return (self, i) if DX_KIND == "f16x2" else None

    def note_rowmax(self, i, grad, rowmax):
        self.rowmax[i] = None if rowmax is None else (grad.data_ptr(), tuple(grad.shape), rowmax)

    def rowmaxes_for(self, dys):
        """Row maxima of the three output gradients if every one of them was noted for exactly these tensors, else None."""
        out = []
        for i, dy in enumerate(dys):
            note = self.rowmax.get(i)
            if note is None or note[0] != dy.data_ptr() or note[1] != tuple(dy.shape):
                return None
            out.append(note[2])
        return out

    def bind(self, x_raw, mean, rstd, gamma, beta, eps, weights, biases, y, xl):
        self.args = (x_raw, mean, rstd, gamma, beta, eps, weights, biases)
        self.n_out = weights[0].shape[0]
        if not self.recompute:
            self.y, self.xl = y, xl

    def get_y(self):
        """relu(LN(x_raw)) for the weight gradients: never kept when the forward GEMM normalises on the fly."""
        if self.y is None:
            x_raw, mean, rstd, gamma, beta, eps = self.args[:6]
            with torch.no_grad():
                self.y = _ln_relu_forward(x_raw, gamma, beta, eps)[1]
        return self.y

    def get_xl(self, g):
        if self.xl is None:
            x_raw, _, _, gamma, beta, eps, weights, biases = self.args
            w2, b2 = [w.detach() for w in weights[:2]], [b.detach() for b in biases[:2]]
            with torch.no_grad():
                if self.ln_fused and self.y is None and gemm_f16x2_ln_y_supported(x_raw.shape[0], self.n_out, x_raw.shape[1],
                                                                                  x_raw.shape[1], 2 * self.n_out):
                    # one pass over x_raw rebuilds the attention sources AND relu(LN(x_raw)) for the weight gradients
                    self.xl, _, _, _, self.y = gemm_f16x2_groups_ln(x_raw, gamma, beta, eps, w2, b2, want_y=True)
                elif self.ln_fused and self.y is None:
                    self.xl = gemm_f16x2_groups_ln(x_raw, gamma, beta, eps, w2, b2)[0]
                else:
                    self.xl = gemm_f16x2_groups(self.get_y(), w2, b2)
        return self.xl[:, g * self.n_out:(g + 1) * self.n_out]

    def xl_getter(self, g):
        return (lambda: self.get_xl(g)) if self.recompute else None

    def release(self):
        self.y = self.xl = None
        self.rowmax = {}


def edge_block_supported(x_raw, gamma, weights_and_biases):
    """The fused front end needs the grouped fp16 GEMM (three projections of equal width, 64 <= d <= 256) and biases."""
    if GEMM_KIND != "f16x2" or WGRAD_KIND != "f16x2" or gamma is None or len(weights_and_biases) != 3:
        return False
    M, K = x_raw.shape
    N = weights_and_biases[0][0].shape[0]
    return (x_raw.is_cuda and M >= TENSOR_CORE_MIN_ROWS and all(w.shape == (N, K) and b is not None for w, b in weights_and_biases)
            and gemm_f16x2_supported(M, N, K, K, 3 * N) and gemm_tf32x3_supported(M, K, 3 * N, N, K)
            and wgrad_f16x2_supported(M, N, K, N, K))


class _EdgeBlockProject(torch.autograd.Function):
    """(x_raw, LayerNorm, W_sp, W_v, W_proj) -> (lin_l_sp(y), lin_l_v(y), lin_proj(y), x_raw) with y = relu(LN(x_raw)).

    Backward is one concatenated input-gradient GEMM, three fp16 weight gradients and the LayerNorm+ReLU backward with
    the residual gradient added in the same pass.  With ``rc.recompute`` nothing of size [E, d] except ``x_raw`` is kept:
    y and the two attention sources are rebuilt when the first backward node of the block asks for them."""

    @staticmethod
    def forward(ctx, x_raw, gamma, beta, eps, rc, *wb):
        weights, biases = wb[0::2], wb[1::2]
        rc.ln_fused = LN_FUSED and (LN_FUSED_ANY_SHAPE or gemm_f16x2_ln_y_supported(
            x_raw.shape[0], weights[0].shape[0], x_raw.shape[1], x_raw.shape[1], 3 * weights[0].shape[0]))
        if rc.ln_fused:
            # LayerNorm + ReLU inside the GEMM's operand producer: x_raw is read once; relu(LN(x_raw)) is written out only when
            # backward will need it and it is not going to be recomputed (same pass, from the registers that hold it)
            x_raw, y = x_raw.contiguous(), None
            gamma, beta = gamma.contiguous(), beta.contiguous()
            M, K = x_raw.shape
            keep_y = ((not rc.recompute) and any(ctx.needs_input_grad) and       # (grad mode is always off inside Function.forward)
                      gemm_f16x2_ln_y_supported(M, weights[0].shape[0], K, K, 3 * weights[0].shape[0]))
            if keep_y:
                out, x_amax, mean, rstd, y = gemm_f16x2_groups_ln(x_raw, gamma, beta, eps, weights, biases, want_y=True)
            else:
                out, x_amax, mean, rstd = gemm_f16x2_groups_ln(x_raw, gamma, beta, eps, weights, biases)
        else:
            x_raw, y, mean, rstd, gamma, beta = _ln_relu_forward(x_raw, gamma, beta, eps)
            out, x_amax = gemm_f16x2_groups(y, weights, biases, want_amax=True)
        N = weights[0].shape[0]
        rc.bind(x_raw, mean, rstd, gamma, beta, eps, weights, biases, y, out)
        ctx.save_for_backward(x_raw, mean, rstd, gamma, beta, x_amax, *weights)
        ctx.rc = rc
        ctx.set_materialize_grads(False)
        return (*(out[:, g * N:(g + 1) * N] for g in range(3)), x_raw.view_as(x_raw))

    @staticmethod
    def backward(ctx, *grads):
        x_raw, mean, rstd, gamma, beta, x_amax, *weights = ctx.saved_tensors
        *dys, dskip = grads
        rc = ctx.rc
        N = weights[0].shape[0]
        if all(dy is None for dy in dys):
            zeros = torch.zeros_like(gamma)
            return (dskip, zeros, zeros, None, None, *([None] * 6))
        y = rc.get_y()
        dys = [torch.zeros((x_raw.shape[0], N), dtype=torch.float32, device=x_raw.device) if dy is None else dy.contiguous()
               for dy in dys]
        wcat = torch.cat([w.t() for w in weights], dim=1)
        rowmaxes = rc.rowmaxes_for(dys) if DX_KIND == "f16x2" else None
        if rowmaxes is not None and gemm_f16x2_cat_supported(x_raw.shape[0], wcat.shape[0], len(dys), N) and (len(dys) * (N // 64)) % 4 == 0:
            # every output gradient came with its row maxima (attention backward kernels, the update's x0 kernel):
            # one-pass fp16 input gradient at twice the 3xTF32 tensor rate
            dy_x, dy_amax = gemm_f16x2_cat(dys, wcat, want_amax=True, rowmax=rowmaxes)
        else:
            dy_x, dy_amax = gemm_tf32x3_cat(dys, wcat, want_amax=True)
        wgrads = []
        if WGRAD_MULTI:
            dw_all, db_all = wgrad_f16x2_multi(dys, y, dy_amax, x_amax)
            for i in range(len(dys)):
                wgrads += [dw_all[i], db_all[i]]
        else:
            for i, dy in enumerate(dys):
                dw, db = wgrad_f16x2(dy, y, dy_amax[i:i + 1], x_amax, with_bias=True)
                wgrads += [dw, db]
        del y
        rc.release()
        dx, dgamma, dbeta = _ln_relu_backward(x_raw, mean, rstd, gamma, beta, dy_x, add=dskip)
        return (dx, dgamma, dbeta, None, None, *wgrads)


def edge_block_project(x_raw, gamma, beta, eps, rc, weights_and_biases):
    flat = [t for wb in weights_and_biases for t in wb]
    return _EdgeBlockProject.apply(x_raw, gamma, beta, eps, rc, *flat)


UPDATE_BWD_FUSED = os.environ.get("GASFM_UPDATE_BWD_FUSED", "1") != "0"    # A/B switch


def _update_backward_views(d_out, x0, W0, scale, plan, want_rowmax=False):
    """The observation update's backward in one storage-order pass over d_out [E,w]: dV[m,w] = scale * per-view sums,
    dx0[E,d0] = scale * d_out W0, dW0[w,d0] = scale * d_out^T x0 and (optionally) max |d_out[e,:]| per row."""
    E, w = d_out.shape
    d0 = x0.shape[1]
    dev = d_out.device
    dV = torch.empty((plan.n_seg, w), dtype=torch.float32, device=dev)
    dx0 = torch.empty((E, d0), dtype=torch.float32, device=dev)
    dW0 = torch.empty((w, d0), dtype=torch.float32, device=dev)
    rowmax = torch.empty(E, dtype=torch.float32, device=dev) if want_rowmax else None
    ws = torch.empty(max(1, _lib.size_query("gasfm_update_bwd_views_ws_bytes", plan.max_chunks, w) // 4), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_update_bwd_views", _lib.ptr(d_out), E, w, _lib.ptr(x0), _lib.ptr(W0), d0, float(scale),
                  _lib.ptr(plan.seg_ptr), plan.n_seg, plan.chunk, _lib.ptr(plan.chunk_ptr), _lib.ptr(plan.chunk_seg), plan.max_chunks,
                  _lib.ptr(dV), _lib.ptr(dx0), _lib.ptr(dW0), _lib.ptr(rowmax), _lib.ptr(ws), _lib.stream_ptr())
    return dV, dx0, dW0, rowmax


def _x0_backward(d_out, x0, W0, scale, want_rowmax=False):
    """-> (dx0, dW0, rowmax | None); ``want_rowmax``: max |d_out[e, :]| per row, from the same pass over d_out."""
    E, w = d_out.shape
    d0 = x0.shape[1]
    dev = d_out.device
    dx0 = torch.empty((E, d0), dtype=torch.float32, device=dev)
    dW0 = torch.empty((w, d0), dtype=torch.float32, device=dev)
    ws = torch.empty(max(1, _lib.size_query("gasfm_x0_bwd_ws_bytes", E, w) // 4), dtype=torch.float32, device=dev)
    rowmax = torch.empty(E, dtype=torch.float32, device=dev) if want_rowmax else None
    with _lib.device_guard(dev):
        if want_rowmax:
            _lib.call("gasfm_x0_bwd_rowmax", _lib.ptr(d_out), E, w, _lib.ptr(x0), _lib.ptr(W0), d0, float(scale),
                      _lib.ptr(dx0), _lib.ptr(dW0), _lib.ptr(ws), _lib.ptr(rowmax), _lib.stream_ptr())
        else:
            _lib.call("gasfm_x0_bwd", _lib.ptr(d_out), E, w, _lib.ptr(x0), _lib.ptr(W0), d0, float(scale),
                      _lib.ptr(dx0), _lib.ptr(dW0), _lib.ptr(ws), _lib.stream_ptr())
    return dx0, dW0, rowmax


class _LinearTinyK(torch.autograd.Function):
    """y = x W^T + b for an input that is only 1..4 wide (the first block reads the 2-d observations):
    write-bound, so one pass over y forward (edge_update kernel without the projected term) and one
    pass over dY backward (the x0 kernel) instead of three K=2 SGEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight = x.contiguous(), weight.contiguous()
        E, d0 = x.shape
        w = weight.shape[0]
        out = torch.empty((E, w), dtype=torch.float32, device=x.device)
        with _lib.device_guard(x.device):
            _lib.call("gasfm_edge_update_fwd", None, w, _lib.ptr(x), d0, _lib.ptr(weight), None, None,
                      _lib.ptr(None if bias is None else bias.contiguous()), None, w, None, None, E, w,
                      0.0, 1.0, _lib.ptr(out), _lib.stream_ptr())
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx, dw, _ = _x0_backward(dy, x, weight, 1.0)
        db = col_sum(dy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


TENSOR_CORE_MIN_ROWS = 4096   # below this the launch overhead dominates; cuBLAS is fine


def linear_multi(x, weights_and_biases):
    """[(W_i, b_i)] -> tuple of x W_i^T + b_i; one autograd node when every shape fits the tensor-core kernel."""
    M, K = x.shape
    lda = x.stride(0) if x.stride(1) == 1 else K
    ok = x.is_cuda and M >= TENSOR_CORE_MIN_ROWS and all(
        b is not None and gemm_tf32x3_supported(M, w.shape[0], K, lda, w.shape[0]) and
        gemm_tf32x3_supported(M, K, w.shape[0], w.shape[0], K) for w, b in weights_and_biases)
    if ok:
        flat = [t for wb in weights_and_biases for t in wb]
        return _LinearMulti.apply(x, *flat)
    return tuple(linear(x, w, b) for w, b in weights_and_biases)


def linear(x, weight, bias=None):
    """``F.linear`` for the observation-level projections: tcgen05 3xTF32 GEMM when the shape allows
    (N <= 256, multiples of 16 / 4), cuBLAS fp32 otherwise (tiny first-layer widths)."""
    M, K = x.shape
    N = weight.shape[0]
    lda = x.stride(0) if x.stride(1) == 1 else K
    if x.is_cuda and M >= TENSOR_CORE_MIN_ROWS and gemm_tf32x3_supported(M, N, K, lda, N):
        return _LinearTC.apply(x, weight, bias)
    if x.is_cuda and x.dtype == torch.float32 and M >= TENSOR_CORE_MIN_ROWS and 1 <= K <= 4 and N % 4 == 0 and N <= 1024:
        return _LinearTinyK.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)
