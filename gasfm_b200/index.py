"""Per-scene observation index: CSR over views, CSC over tracks, chunk tables (device resident).

The reference re-derives its graph structure on every call: ``to_torch_hybrid_sparse_coo()
.coalesce()`` sorts the int64 indices ~49 times per forward (``code/utils/sparse_utils.py:436-449``,
``code/models/layers.py:545-549,561-565,823,922``) and every GAT call re-checks
``torch.all(valid_indices == M.indices())`` with a host sync (``code/utils/dataset_utils.py:557``).
Here the structure is built ONCE per scene by ``gasfm_csr_build`` / ``gasfm_plan_chunks`` and
cached on the scene object; the kernels only read int32 arrays.
"""
import torch

from . import _lib

MIN_N_VIEWS_PER_POINT = 2   # reference: code/utils/constants.py:2
MIN_N_POINTS_PER_VIEW = 8   # reference: code/utils/constants.py:6


def _pow2ceil(x):
    p = 1
    while p < x:
        p <<= 1
    return p


class _DeferredValidation:
    """While active, ``ObservationIndex`` does not read its status word back (no host synchronisation, so the build can be
    captured into a CUDA graph); the status tensors are collected here and checked later with ``check``."""
    active = None

    def __init__(self):
        self.entries = []          # (status tensor, m, n)

    def __enter__(self):
        self._outer, _DeferredValidation.active = _DeferredValidation.active, self
        return self

    def __exit__(self, *exc):
        _DeferredValidation.active = self._outer
        return False


def deferred_validation():
    return _DeferredValidation()


def raise_for_status(st, m, n):
    if st & 1:
        raise ValueError("observation indices out of range for an (m=%d, n=%d) scene" % (m, n))
    if st & 2:
        raise ValueError("observation indices must be row-major sorted without duplicates "
                         "(the order np.nonzero / coalesce() produce)")


def one_segment_ptr(k, device):
    """int32 [0, k] built by device fills (``torch.tensor([0, k], device=...)`` is a pageable host copy, which a stream
    capture refuses)."""
    seg_ptr = torch.full((2,), int(k), dtype=torch.int32, device=device)
    seg_ptr[:1].zero_()
    return seg_ptr


class SegmentPlan:
    """One aggregation graph as segments over observation rows (see include/gasfm_b200.h)."""

    def __init__(self, seg_ptr, perm, n_seg, n_edges, chunk, device):
        self.seg_ptr = seg_ptr            # int32 [T+1]
        self.perm = perm                  # int32 [E'] or None (storage order)
        self.n_seg = int(n_seg)
        self.n_edges = int(n_edges)       # number of edges the plan covers
        self.chunk = int(chunk)
        self.chunk_ptr = None
        self.chunk_seg = None
        self.max_chunks = 0
        self._ws = {}
        if self.chunk > 0 and self.n_seg > 0:
            self.max_chunks = self.n_edges // self.chunk + self.n_seg + 1
            self.chunk_ptr = torch.empty(self.n_seg + 1, dtype=torch.int32, device=device)
            self.chunk_seg = torch.zeros(self.max_chunks, dtype=torch.int32, device=device)
            _lib.call("gasfm_plan_chunks", _lib.ptr(seg_ptr), self.n_seg, self.chunk, _lib.ptr(self.chunk_ptr),
                      _lib.ptr(self.chunk_seg), self.max_chunks, _lib.stream_ptr())

    def workspace(self, nbytes, device):
        """fp32 scratch of at least ``nbytes`` bytes, shared by every kernel that runs over this plan (attention
        forward / backward, segment sums, the loss).  SINGLE-STREAM: callers on different streams must use different
        plans (the track-sharded tests give every rank its own scene, hence its own plans)."""
        n = (nbytes + 3) // 4
        buf = self._ws.get("buf")
        if buf is None or buf.numel() < n or buf.device != device:
            if buf is not None:
                # never hand the old buffer back to the allocator: a captured CUDA graph (gasfm_b200.graphs) may have
                # its address baked in and would otherwise write into freed / reused memory on its next replay
                self._ws.setdefault("retired", []).append(buf)
            buf = torch.empty(max(n, 1), dtype=torch.float32, device=device)
            self._ws["buf"] = buf
        return buf

    def abi_args(self):
        """(seg_ptr, perm, n_seg, chunk, chunk_ptr, chunk_seg, max_chunks) in C-ABI order."""
        return (_lib.ptr(self.seg_ptr), _lib.ptr(self.perm), self.n_seg, self.chunk,
                _lib.ptr(self.chunk_ptr), _lib.ptr(self.chunk_seg), self.max_chunks)


def csr_chunk(n_edges):
    """Chunk length for multi-segment (view) plans: enough warps to fill 148 SMs, <= 128 edges."""
    return max(8, min(128, _pow2ceil(max(1, n_edges // 4096))))


def single_segment_chunk(n_edges):
    """Chunk length for one-segment (global) plans: ~1k chunks, merged by one CTA."""
    return max(8, min(1024, _pow2ceil(max(1, n_edges // 1024))))


class ObservationIndex:
    """CSR/CSC + plans for one scene, built from the reference-format ``indices[2,E]`` (int64)."""

    def __init__(self, indices, m, n, validate=True):
        if not indices.is_cuda:
            raise RuntimeError("gasfm_b200: the observation index is built on the GPU; move the scene to a CUDA "
                               "device first (there is no CPU path)")
        if indices.dtype != torch.int64 or indices.dim() != 2 or indices.shape[0] != 2:
            raise ValueError("indices must be an int64 tensor of shape [2, E]")
        indices = indices.contiguous()
        dev = indices.device
        E = indices.shape[1]
        self.m, self.n, self.n_obs, self.device = int(m), int(n), int(E), dev
        i32 = dict(dtype=torch.int32, device=dev)
        self.row_idx = torch.empty(E, **i32)
        self.col_idx = torch.empty(E, **i32)
        self.row_ptr = torch.empty(m + 1, **i32)
        self.col_ptr = torch.empty(n + 1, **i32)
        self.csc_perm = torch.empty(E, **i32)
        status = torch.zeros(1, **i32)
        ws = torch.empty(_lib.size_query("gasfm_csr_build_ws_bytes", E, self.n) // 4 + 1, **i32)
        with _lib.device_guard(dev):
            _lib.call("gasfm_csr_build", _lib.ptr(indices), E, self.m, self.n, _lib.ptr(self.row_idx),
                      _lib.ptr(self.col_idx), _lib.ptr(self.row_ptr), _lib.ptr(self.col_ptr),
                      _lib.ptr(self.csc_perm), _lib.ptr(status), _lib.ptr(ws), _lib.stream_ptr())
            self.status = status
            if _DeferredValidation.active is not None:
                _DeferredValidation.active.entries.append((status, self.m, self.n))
            elif validate:
                raise_for_status(int(status.item()), self.m, self.n)
            # views: long contiguous segments -> chunked; tracks: short gathered segments
            self.by_view = SegmentPlan(self.row_ptr, None, self.m, E, csr_chunk(E), dev)
            self.by_track = SegmentPlan(self.col_ptr, self.csc_perm, self.n, E, 0, dev)
        self._global_plans = {}

    # -- the two single-target graphs of ViewAndScenePoint2Global -------------------------------
    def global_plan(self, kind, valid_ids):
        """Plan for view2global / scenepoint2global: one segment holding the valid views
        (>= 8 points) or tracks (>= 2 views), ``code/datasets/SceneData.py:174-187``.
        ``valid_ids`` int64 [k] (sorted ids of the valid views / tracks)."""
        key = (kind, int(valid_ids.data_ptr()), int(valid_ids.numel()))
        plan = self._global_plans.get(key)
        if plan is None:
            dev = self.device
            k = int(valid_ids.numel())
            total = self.m if kind == "view" else self.n
            seg_ptr = one_segment_ptr(k, dev)
            perm = None if k == total else valid_ids.to(torch.int32)
            with _lib.device_guard(dev):
                plan = SegmentPlan(seg_ptr, perm, 1, k, single_segment_chunk(k), dev)
            plan.covers_all = perm is None
            self._global_plans[key] = plan
        return plan


def plan_from_targets(dst, n_targets):
    """Generic plan for an arbitrary target list ``dst[E]`` (int64, any order): stable grouping of
    edge ids by target.  Used by ``GATv2Conv.forward(x, edge_index)`` for graphs that are not one
    of the scene's cached plans."""
    dev = dst.device
    E = int(dst.numel())
    pairs = torch.stack((torch.arange(E, dtype=torch.int64, device=dev), dst.to(torch.int64))).contiguous()
    i32 = dict(dtype=torch.int32, device=dev)
    row_idx, col_idx = torch.empty(E, **i32), torch.empty(E, **i32)
    row_ptr, col_ptr = torch.empty(E + 1, **i32), torch.empty(n_targets + 1, **i32)
    perm, status = torch.empty(E, **i32), torch.zeros(1, **i32)
    ws = torch.empty(_lib.size_query("gasfm_csr_build_ws_bytes", E, int(n_targets)) // 4 + 1, **i32)
    with _lib.device_guard(dev):
        _lib.call("gasfm_csr_build", _lib.ptr(pairs), E, max(E, 1), int(n_targets), _lib.ptr(row_idx), _lib.ptr(col_idx),
                  _lib.ptr(row_ptr), _lib.ptr(col_ptr), _lib.ptr(perm), _lib.ptr(status), _lib.ptr(ws), _lib.stream_ptr())
        if int(status.item()) != 0:
            raise ValueError("edge targets out of range")
        avg = E / max(1, n_targets)
        chunk = 0 if avg <= 64 else (single_segment_chunk(E) if n_targets == 1 else csr_chunk(E))
        return SegmentPlan(col_ptr, perm, n_targets, E, chunk, dev)


_INDEX_ATTR = "_gasfm_b200_index"


def index_for(x):
    """ObservationIndex of a SparseMat-like object (``.indices [2,E]``, ``.shape (m,n,d)``),
    cached on the object and shared by every SparseMat derived from it."""
    idx = getattr(x, _INDEX_ATTR, None)
    if idx is not None and idx.device == x.indices.device and idx.n_obs == x.indices.shape[1]:
        return idx
    idx = ObservationIndex(x.indices, x.shape[0], x.shape[1])
    try:
        setattr(x, _INDEX_ATTR, idx)
    except AttributeError:
        pass
    return idx
