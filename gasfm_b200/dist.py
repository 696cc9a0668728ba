"""Track-sharded multi-GPU execution of the GASFM attention path (one process per GPU).

The reference has no distributed code at all (single ``cuda:0``, ``code/main.py:78``); this is the
B200 design of SURVEY.md section 8(e).

Partition.  Tracks (columns of the observation matrix) are split into ``world`` contiguous blocks of
equal *edge* load, so every observation of a point lives on exactly one GPU: the column direction
(proj2scenepoint, the per-observation update, all point-level layers) needs no communication.  View
features ``[m, .]``, the global feature and all parameters are replicated, and every replicated
quantity is computed bit-identically on every rank.

Forward exchange.  A view's softmax runs over observations held by all ranks.  Each rank computes, with
the ordinary fused kernel, the un-normalised partial ``(max_g, sum_g, acc_g)`` over its local edges; ONE
kernel per aggregation (``gasfm_peer_lse_merge``) pushes the partial into every peer's exchange buffer
over NVLink, waits for the peers' flags and merges like flash-attention blocks, in rank order:
    M = max_g max_g;   [L, A] = sum_g e^{max_g-M} [sum_g, acc_g];   out = A / L + bias
The same merge (with one target) serves scenepoint2global.

Backward.  Replicated tensors carry FULL gradients, identical on every rank: the loss terms computed from
replicated predictions (``Ps_norm``) are evaluated on every rank, the terms of local predictions
(``pts3D`` of the rank's tracks) locally.  Gradients are summed exactly where a replicated tensor feeds
local work -- the per-view query ``XR`` of the sharded aggregation, the per-view / global terms of the
observation update (``V[row]``, ``g``; ``code/models/layers.py:941-945``) -- by one
``gasfm_peer_allreduce_sum`` each (``[m, d]`` floats).  The output gradient of a merge is already full, so
the backward of a merge needs no exchange at all.  Parameters applied to replicated tensors (the 145 M
view-/global-level weights) therefore hold complete gradients without communication; only the
observation-/point-level parameters (``is_local_parameter``; ~1 % of the model) hold partial sums, and
one exchange over a persistent flat bucket (``LocalGradBucket``) completes them.

No NCCL call is on the step's path (torch.distributed only carries the IPC handles at set-up), so a
whole sharded step -- forward, loss, backward, exchanges -- is captured in one CUDA graph
(``gasfm_b200.graphs.GraphedStep``).  ``CollectiveExchange`` is the library arm (all_gather through
torch.distributed + the same merge kernels on the gathered buffer): the A/B baseline, and what the
single-GPU multi-process tests use.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .scene import Scene
from .utils.constants import MIN_N_POINTS_PER_VIEW
from .utils.dataset_utils import AxialAggregationGraphWrapper


# ---------------------------------------------------------------------------------------------
# host logic: partition
# ---------------------------------------------------------------------------------------------
def partition_tracks(views_per_track, world):
    """Boundaries ``b[0..world]`` of contiguous track blocks with (nearly) equal edge counts:
    rank g owns tracks [b[g], b[g+1]).  Balances sum_j k_j, not n/world."""
    k = np.asarray(views_per_track, dtype=np.int64)
    n = k.size
    cum = np.concatenate(([0], np.cumsum(k)))
    total = cum[-1]
    bounds = [0]
    for g in range(1, world):
        target = total * g / world
        j = int(np.searchsorted(cum, target, side="left"))
        if j > 0 and abs(cum[j - 1] - target) <= abs(cum[min(j, n)] - target):
            j -= 1
        bounds.append(min(max(j, bounds[-1]), n))
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def shard_observations(indices, values, m, n, rank, world, bounds=None):
    """Local slice of a row-major sorted observation list: the observations of this rank's tracks,
    with local column ids.  Row-major order is preserved, so the local list is itself a valid scene."""
    indices = np.asarray(indices)
    if bounds is None:
        bounds = partition_tracks(np.bincount(indices[1], minlength=n), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    sel = (indices[1] >= lo) & (indices[1] < hi)
    local = np.stack((indices[0][sel], indices[1][sel] - lo))
    return local, np.asarray(values)[sel], lo, hi


class ShardInfo:
    """Which tracks this rank holds, and the exchange that connects it to the other ranks."""

    def __init__(self, rank, world, col_begin, col_end, n_global, exchange=None):
        self.rank, self.world = rank, world
        self.col_begin, self.col_end, self.n_global = col_begin, col_end, n_global
        self.exchange = exchange


def shard_scene(indices, values, m, n, rank, world, exchange=None, bounds=None):
    """Scene holding this rank's tracks.  ``view2global`` uses the GLOBAL per-view counts (views with
    >= 8 points anywhere, ``code/datasets/SceneData.py:174``); the view-aggregation and
    scenepoint2global graphs are flagged so that the model merges their partials through ``exchange``."""
    indices = np.asarray(indices)
    local_idx, local_vals, lo, hi = shard_observations(indices, values, m, n, rank, world, bounds)
    if local_idx.shape[1] == 0:
        raise ValueError(f"rank {rank} of {world} would hold no observations ({n} tracks): use fewer ranks for this scene")
    scene = Scene.from_observations(local_idx, local_vals, m, hi - lo)
    pts_per_view = np.bincount(indices[0], minlength=m)
    scene.x.pts_per_cam = torch.from_numpy(pts_per_view).unsqueeze(1)       # global counts (replicated)
    rows = torch.from_numpy(np.nonzero(pts_per_view >= MIN_N_POINTS_PER_VIEW)[0])
    scene.graph_wrappers["view2global"] = AxialAggregationGraphWrapper(
        m, 1, 0, valid_indices=torch.stack((rows, torch.zeros_like(rows))))
    scene.shard = ShardInfo(rank, world, lo, hi, n, exchange)
    scene.x.shard = scene.shard
    scene.graph_wrappers["proj2view"].shard = scene.shard
    scene.graph_wrappers["scenepoint2global"].shard = scene.shard
    return scene


# ---------------------------------------------------------------------------------------------
# exchanges
# ---------------------------------------------------------------------------------------------
def _pad4(n):
    return (int(n) + 3) // 4 * 4


class _ExchangeBase:
    """Shared front end: ``lse_merge`` and ``allreduce_sum`` on fp32 CUDA tensors."""

    world = 1
    region_floats = 0

    def lse_merge(self, acc, seg_max, seg_sum, heads, bias=None):
        """Merge per-rank un-normalised softmax partials.  acc [T,H*C] = sum_e e^{s-max} x_e, seg_max / seg_sum
        [T,H] -> (normalised out [T,H*C] (+ bias), M [T,H], L [T,H]), bit-identical on every rank."""
        acc, seg_max, seg_sum = acc.contiguous(), seg_max.contiguous(), seg_sum.contiguous()
        T, hc = acc.shape
        out = torch.empty_like(acc)
        M, L = torch.empty_like(seg_max), torch.empty_like(seg_sum)
        self._lse(acc, seg_max, seg_sum, None if bias is None else bias.contiguous(), T, heads, hc // heads, out, M, L)
        return out, M, L

    def allreduce_sum(self, t, out=None, scale=1.0):
        """Sum of ``t`` over the ranks (same value on every rank; summed in rank order).  ``out`` may be ``t``."""
        src = t.contiguous()
        n = src.numel()
        flat = src.view(-1)
        if n % 4:                                   # the kernels move float4: pad odd sizes
            flat = torch.cat((flat, flat.new_zeros(_pad4(n) - n)))
        dst = flat if (out is t and flat.data_ptr() == t.data_ptr()) else torch.empty_like(flat)
        for off in range(0, flat.numel(), self.region_floats):
            cnt = min(self.region_floats, flat.numel() - off)
            self._sum(flat[off:off + cnt], dst[off:off + cnt], cnt, float(scale))
        res = dst[:n].view(t.shape)
        if out is not None and out.data_ptr() != res.data_ptr():
            out.copy_(res)
            return out
        return res

    def barrier(self):
        pass

    def check(self):
        """Raise if an exchange timed out (synchronises)."""


class PeerExchange(_ExchangeBase):
    """The product path: hand-written exchange kernels over NVLink peer memory (``csrc/peer_comm.cu``).

    ``torch.distributed`` (any backend) is used ONCE, to pass the IPC handles of the exchange buffers around;
    afterwards every exchange is a single kernel launch on the current stream."""

    def __init__(self, device, group=None, region_floats=1 << 22, timeout_s=20.0):
        self.device = torch.device(device)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.region_floats = _pad4(region_floats)
        lib = _lib.load()
        with _lib.device_guard(self.device):
            self._buf, self._flags = ctypes.c_void_p(), ctypes.c_void_p()
            _lib.call_setup("gasfm_peer_alloc", lib.gasfm_peer_buffer_bytes(self.world, self.region_floats), ctypes.byref(self._buf))
            _lib.call_setup("gasfm_peer_alloc", lib.gasfm_peer_flags_bytes(self.world), ctypes.byref(self._flags))
            mine = []
            for p in (self._buf, self._flags):
                h = ctypes.create_string_buffer(64)
                _lib.call_setup("gasfm_peer_export", p, h)
                mine.append(h.raw)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            self._imported = []
            bufs, flags = (ctypes.c_void_p * self.world)(), (ctypes.c_void_p * self.world)()
            for r, (hb, hf) in enumerate(handles):
                if r == self.rank:
                    bufs[r], flags[r] = self._buf.value, self._flags.value
                    continue
                pb, pf = ctypes.c_void_p(), ctypes.c_void_p()
                _lib.call_setup("gasfm_peer_import", ctypes.create_string_buffer(hb, 64), ctypes.byref(pb))
                _lib.call_setup("gasfm_peer_import", ctypes.create_string_buffer(hf, 64), ctypes.byref(pf))
                self._imported += [pb, pf]
                bufs[r], flags[r] = pb.value, pf.value
            self._comm = ctypes.c_void_p()
            _lib.call_setup("gasfm_peer_comm_create", self.rank, self.world, bufs, flags, self.region_floats,
                            float(timeout_s), ctypes.byref(self._comm))
        dist.barrier(group=group)        # every rank has imported every buffer before the first push

    @classmethod
    def local_group(cls, world, device, region_floats=1 << 20, timeout_s=20.0):
        """``world`` communicators inside ONE process (same-device buffers, no IPC): the ranks run on different
        streams / threads of one GPU.  The kernels are the multi-GPU ones; used by the single-GPU tests."""
        device = torch.device(device)
        lib = _lib.load()
        region = _pad4(region_floats)
        ptrs = []
        with _lib.device_guard(device):
            for _ in range(world):
                b, f = ctypes.c_void_p(), ctypes.c_void_p()
                _lib.call_setup("gasfm_peer_alloc", lib.gasfm_peer_buffer_bytes(world, region), ctypes.byref(b))
                _lib.call_setup("gasfm_peer_alloc", lib.gasfm_peer_flags_bytes(world), ctypes.byref(f))
                ptrs.append((b, f))
            group = []
            for r in range(world):
                ex = cls.__new__(cls)
                ex.device, ex.group, ex.rank, ex.world, ex.region_floats = device, None, r, world, region
                ex._buf, ex._flags = ptrs[r]
                ex._imported = []
                bufs, flags = (ctypes.c_void_p * world)(), (ctypes.c_void_p * world)()
                for q in range(world):
                    bufs[q], flags[q] = ptrs[q][0].value, ptrs[q][1].value
                ex._comm = ctypes.c_void_p()
                _lib.call_setup("gasfm_peer_comm_create", r, world, bufs, flags, region, float(timeout_s), ctypes.byref(ex._comm))
                group.append(ex)
        return group

    def _lse(self, acc, mx, sm, bias, T, H, C, out, M, L):
        with _lib.device_guard(acc.device):
            _lib.call("gasfm_peer_lse_merge", self._comm, _lib.ptr(acc), _lib.ptr(mx), _lib.ptr(sm), _lib.ptr(bias),
                      T, H, C, _lib.ptr(out), _lib.ptr(M), _lib.ptr(L), _lib.stream_ptr())

    def _sum(self, src, dst, n, scale):
        with _lib.device_guard(src.device):
            _lib.call("gasfm_peer_allreduce_sum", self._comm, _lib.ptr(src), _lib.ptr(dst), n, scale, _lib.stream_ptr())

    def barrier(self):
        with _lib.device_guard(self.device):
            _lib.call("gasfm_peer_allreduce_sum", self._comm, None, None, 0, 1.0, _lib.stream_ptr())

    def check(self):
        err = ctypes.c_int(0)
        with _lib.device_guard(self.device):
            _lib.call_setup("gasfm_peer_comm_error", self._comm, ctypes.byref(err))
        if err.value:
            raise RuntimeError("gasfm_b200: a peer-memory exchange timed out (a rank did not reach the matching exchange)")


class CollectiveExchange(_ExchangeBase):
    """Library arm: ``torch.distributed.all_gather`` (NCCL, or gloo with CUDA tensors) of every rank's partial, then
    the same merge kernels on the gathered buffer.  Two launches + a collective per exchange; not graph-capturable."""

    def __init__(self, device, group=None, region_floats=1 << 22):
        self.device = torch.device(device)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.region_floats = _pad4(region_floats)

    def _gather(self, packed):
        parts = [torch.empty_like(packed) for _ in range(self.world)]
        dist.all_gather(parts, packed, group=self.group)
        return torch.stack(parts).contiguous()

    def _lse(self, acc, mx, sm, bias, T, H, C, out, M, L):
        packed = torch.cat((acc.reshape(-1), mx.reshape(-1), sm.reshape(-1)))
        region = _pad4(packed.numel())
        if region != packed.numel():
            packed = torch.cat((packed, packed.new_zeros(region - packed.numel())))
        gathered = self._gather(packed)
        with _lib.device_guard(acc.device):
            _lib.call("gasfm_lse_merge_gathered", _lib.ptr(gathered), self.world, region, _lib.ptr(bias), T, H, C,
                      _lib.ptr(out), _lib.ptr(M), _lib.ptr(L), _lib.stream_ptr())

    def _sum(self, src, dst, n, scale):
        gathered = self._gather(src.contiguous())
        with _lib.device_guard(src.device):
            _lib.call("gasfm_sum_gathered", _lib.ptr(gathered), self.world, n, n, scale, _lib.ptr(dst), _lib.stream_ptr())

    def barrier(self):
        dist.barrier(group=self.group)


# ---------------------------------------------------------------------------------------------
# the sharded aggregation
# ---------------------------------------------------------------------------------------------
class CudaEdgeBackend:
    """Local compute of the sharded GAT on the sm_100a kernels."""

    @staticmethod
    def partial(XL, XR, att, plan, heads):
        from . import ops
        return ops.gat_edge_partial(XL, XR, att, plan, heads)

    @staticmethod
    def backward(XL, XR, att, out_nobias, M, L, d_out, plan, heads, want_rowmax=False):
        from . import ops
        return ops.gat_edge_backward_raw(XL, XR, att, out_nobias, M, L, d_out, plan, heads, want_rowmax=want_rowmax)


class _ShardedGat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, XL, XR, att, bias, plan, heads, exchange, backend, lazy_xl, rowmax_slot=None):
        acc, mx, sm = backend.partial(XL, XR, att, plan, heads)
        out, M, L = exchange.lse_merge(acc, mx, sm, heads, bias)
        if lazy_xl is None:
            ctx.save_for_backward(XL, XR, att, bias, out, M, L)
        else:
            ctx.save_for_backward(XR, att, bias, out, M, L)       # XL is recomputed in backward (ops.LayerRecompute)
        ctx.plan, ctx.heads, ctx.exchange, ctx.backend, ctx.lazy_xl = plan, heads, exchange, backend, lazy_xl
        ctx.rowmax_slot = rowmax_slot
        return out

    @staticmethod
    def backward(ctx, d_out):
        from . import ops
        if ctx.lazy_xl is None:
            XL, XR, att, bias, out, M, L = ctx.saved_tensors
        else:
            XR, att, bias, out, M, L = ctx.saved_tensors
            XL = ctx.lazy_xl()
        # d_out is the FULL gradient of the replicated output, identical on every rank: no exchange is needed for it
        d_out = d_out.contiguous()
        d_bias = None if bias is None else ops.col_sum(d_out)
        out_nobias = out if bias is None else out - bias
        if ctx.rowmax_slot is not None:
            dXL, dXR, datt, rowmax = ctx.backend.backward(XL, XR, att, out_nobias, M, L, d_out, ctx.plan, ctx.heads, want_rowmax=True)
            ctx.rowmax_slot[0].note_rowmax(ctx.rowmax_slot[1], dXL, rowmax)
        else:
            dXL, dXR, datt = ctx.backend.backward(XL, XR, att, out_nobias, M, L, d_out, ctx.plan, ctx.heads)
        dXR = ctx.exchange.allreduce_sum(dXR)           # the query is replicated, its edges are spread over the ranks
        return dXL, dXR, datt.view(att.shape), d_bias, None, None, None, None, None, None


def sharded_gat(XL, XR, att, bias, plan, heads, exchange, backend=CudaEdgeBackend, lazy_xl=None, rowmax_slot=None):
    return _ShardedGat.apply(XL, XR, att, bias, plan, heads, exchange, backend, lazy_xl, rowmax_slot)


class _ReplicatedToLocal(torch.autograd.Function):
    """Identity on a replicated tensor that is about to be consumed by rank-local work; backward sums the partial
    gradients of all ranks, so that the replicated producer sees the full gradient."""

    @staticmethod
    def forward(ctx, t, exchange):
        ctx.exchange = exchange
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return ctx.exchange.allreduce_sum(g.contiguous()), None


def replicated_to_local(t, shard):
    if shard is None or shard.world == 1 or shard.exchange is None:
        return t
    return _ReplicatedToLocal.apply(t, shard.exchange)


# ---------------------------------------------------------------------------------------------
# parameter gradients
# ---------------------------------------------------------------------------------------------
_LOCAL_MARKERS = (
    "embed.", "prev_projfeat_norm_layer.", "residual_skipconn_proj_norm_layer.", "skip_projection.",
    ".proj2scenepoint.", ".global2scenepoint.", "proj2view.graph_conv.lin_l.", "proj2view.graph_conv.att",
    "graph_conv_scenepoint2global.lin_l.", "graph_conv_scenepoint2global.att",
    "projection_feature_update.scenepoint_norm_layer.", "projection_feature_update.lin_scenepoint.",
    "projection_feature_update.lin_proj.", "projection_feature_update.mlp.", "scenepoint_head.", "depth_head.",
)
_REPLICATED_IN_LOCAL = (".global2scenepoint.global_norm_layer.", ".global2scenepoint.lin_global.")


def is_local_parameter(name):
    """True for parameters applied to observation- or point-level tensors of a track-sharded scene: their gradient
    on a rank is a partial sum over the rank's tracks.  Everything else acts on replicated tensors (views, global)
    and already holds the complete gradient on every rank."""
    name = "." + name
    if any(k in name for k in _REPLICATED_IN_LOCAL):
        return False
    return any(k in name for k in _LOCAL_MARKERS)


class LocalGradBucket:
    """Persistent flat buffer behind the ``.grad`` of every selected parameter: autograd accumulates into views of it, and one
    in-place exchange per step completes the partial sums (no per-step ``torch.cat``).

    Track-sharded scenes: ``select`` = ``is_local_parameter`` (default), ``exchange`` = the peer-memory exchange.
    Scene-per-GPU training with graph-replayed per-scene steps: ``select=lambda name: True`` and ``allreduce_nccl`` after
    the replays (every gradient is a per-rank partial there)."""

    def __init__(self, model, exchange=None, select=is_local_parameter):
        self.exchange = exchange
        named = [(k, p) for k, p in model.named_parameters() if p.requires_grad]
        self.local = [p for k, p in named if select(k)]
        self.replicated = [p for k, p in named if not select(k)]
        total = _pad4(sum(p.numel() for p in self.local))
        dev = self.local[0].device if self.local else torch.device("cpu")
        self.flat = torch.zeros(max(total, 4), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.local:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self):
        """Point the selected ``.grad``s at the bucket (without zeroing it) and drop the others."""
        for p, v in zip(self.local, self.views):
            p.grad = v
        for p in self.replicated:
            p.grad = None

    def prepare(self):
        """Before forward: zero the bucket, point the local ``.grad``s at it, drop the replicated ones."""
        self.flat.zero_()
        self.attach()

    def allreduce(self):
        """After backward: complete the local gradients (in place) through the peer-memory exchange."""
        self.exchange.allreduce_sum(self.flat, out=self.flat)

    def allreduce_nccl(self, group=None):
        """One NCCL SUM all-reduce over the whole bucket (scene-per-GPU training)."""
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


class BucketedGradReducer:
    """Scene-per-GPU data parallelism (BASELINE.json configs[3]): every rank trains the full model on its own scene(s);
    the batch gradient is the SUM over scenes (``batch_loss += loss``, ``code/train.py:61-88``).  The 580 MB of
    gradients live in a few persistent flat buckets (``.grad`` = views, filled in reverse layer order as backward
    proceeds); each bucket's NCCL all-reduce is launched on a side stream the moment its last gradient has been
    accumulated, so the exchange over NVLink hides behind the rest of backward.

        reducer = BucketedGradReducer(model)
        reducer.prepare(); loss.backward(); reducer.finish(); optimizer.step()

    ``accumulate_only()``: a context in which backward only accumulates (more scenes than ranks: all but the last scene)."""

    def __init__(self, model, group=None, bucket_bytes=64 << 20):
        self.group = group
        params = [p for p in model.parameters() if p.requires_grad][::-1]      # ~ the order gradients become ready
        self.buckets, cur, cur_n = [], [], 0
        for p in params:
            cur.append(p)
            cur_n += p.numel()
            if cur_n * 4 >= bucket_bytes:
                self.buckets.append(cur)
                cur, cur_n = [], 0
        if cur:
            self.buckets.append(cur)
        self.flats, self.views, self.bucket_of = [], {}, {}
        for b, plist in enumerate(self.buckets):
            flat = torch.zeros(_pad4(sum(p.numel() for p in plist)), dtype=torch.float32, device=plist[0].device)
            off = 0
            for p in plist:
                self.views[p] = flat[off:off + p.numel()].view_as(p)
                self.bucket_of[p] = b
                off += p.numel()
            self.flats.append(flat)
        self.comm_stream = torch.cuda.Stream(device=params[0].device) if params and params[0].is_cuda else None
        self.enabled = True
        self._pending, self._works, self._launched = [], [], []
        for p in params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    def prepare(self):
        for flat in self.flats:
            flat.zero_()
        for p, v in self.views.items():
            p.grad = v
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []

    def _launch(self, b):
        self._launched[b] = True
        if self.comm_stream is None:
            self._works.append(dist.all_reduce(self.flats[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self._works.append(dist.all_reduce(self.flats[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p):
        if not self.enabled or not self._pending:
            return
        b = self.bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._launched[b]:
            self._launch(b)

    def accumulate_only(self):
        reducer = self

        class _Ctx:
            def __enter__(self):
                reducer.enabled = False

            def __exit__(self, *exc):
                reducer.enabled = True
        return _Ctx()

    def finish(self):
        """Launch what is still pending (parameters that received no gradient) and make the current stream wait."""
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        for w in self._works:
            w.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._works = []


def allreduce_gradients(parameters, group=None):
    """One flat NCCL all-reduce SUM over ALL parameter gradients: the gradient exchange of scene-per-GPU data
    parallelism (``batch_loss += loss`` over the scenes of a batch, ``code/train.py:61-88``)."""
    params = [p for p in parameters if p.requires_grad]
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = flat[off:off + n].view_as(p)
        off += n
    return flat


def gather_points(pts_local, shard, group=None):
    """All-gather the per-rank ``pts3D [4, n_local]`` blocks into ``[4, n]`` (track order)."""
    sizes = [torch.zeros(1, dtype=torch.int64, device=pts_local.device) for _ in range(shard.world)]
    dist.all_gather(sizes, torch.tensor([pts_local.shape[1]], dtype=torch.int64, device=pts_local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros(pts_local.shape[0], pad, dtype=pts_local.dtype, device=pts_local.device)
    buf[:, : pts_local.shape[1]] = pts_local
    parts = [torch.empty_like(buf) for _ in range(shard.world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:, :s] for p, s in zip(parts, sizes)], dim=1)
