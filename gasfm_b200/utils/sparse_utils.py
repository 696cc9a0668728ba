"""Mirror of the reference's ``code/utils/sparse_utils.py`` for the attention path.

``SparseMat`` keeps the reference's fields (values [E,d] f32, indices [2,E] i64 row-major,
cam_per_pts [n,1], pts_per_cam [m,1], shape) and methods (``sparse_utils.py:392-449``).  Row /
column pooling runs on the segment kernels instead of ``index_add`` / ``torch.sparse.sum``, and a
device-resident CSR/CSC index (``gasfm_b200.index.ObservationIndex``) is cached on the object
and inherited by every SparseMat derived from it, so nothing is ever re-sorted.
"""
import copy

import torch

from .. import ops
from ..index import index_for, _INDEX_ATTR


class SparseMat:
    def __init__(self, values, indices, cam_per_pts, pts_per_cam, shape, _index=None):
        assert len(shape) == 3
        self.values = values
        self.indices = indices
        self.shape = shape
        self.cam_per_pts = cam_per_pts
        self.pts_per_cam = pts_per_cam
        self.device = self.values.device if self.values is not None else self.indices.device
        if _index is not None:
            setattr(self, _INDEX_ATTR, _index)

    @property
    def size(self):
        return self.shape

    @property
    def index(self):
        """CSR/CSC observation index (built on first use, on the GPU)."""
        return index_for(self)

    def with_values(self, values, n_feat=None):
        """Same sparsity pattern (and cached index), new feature values (``None`` + ``n_feat``: a placeholder
        whose values are never materialised, e.g. relu(LN(x)) of a block that fuses it into its projections)."""
        n_feat = values.shape[1] if n_feat is None else n_feat
        return SparseMat(values, self.indices, self.cam_per_pts, self.pts_per_cam,
                         (self.shape[0], self.shape[1], n_feat), _index=getattr(self, _INDEX_ATTR, None))

    def sum(self, dim):
        """Sum over views (dim=0 -> [n,d]) or over tracks (dim=1 -> [m,d]); sparse_utils.py:406-412."""
        assert dim == 1 or dim == 0
        idx = self.index
        if dim == 1:
            return ops.seg_pool(self.values, idx.by_view, idx.row_idx)
        return ops.seg_pool(self.values, idx.by_track, idx.col_idx)

    def mean(self, dim):
        """sum / count, with the reference's 0/0 = nan for empty rows or columns (sparse_utils.py:414-419)."""
        assert dim == 1 or dim == 0
        if dim == 0:
            return self.sum(dim=0) / self.cam_per_pts
        return self.sum(dim=1) / self.pts_per_cam

    def to(self, device, **kwargs):
        ret = copy.copy(self)
        ret.device = device
        ret.values = ret.values.to(device, **kwargs)
        ret.indices = ret.indices.to(device, **kwargs)
        ret.pts_per_cam = ret.pts_per_cam.to(device, **kwargs)
        ret.cam_per_pts = ret.cam_per_pts.to(device, **kwargs)
        idx = getattr(ret, _INDEX_ATTR, None)
        if idx is not None and idx.device != ret.indices.device:
            delattr(ret, _INDEX_ATTR)
        return ret

    def __add__(self, other):
        assert self.shape == other.shape
        return self.with_values(self.values + other.values)

    def to_torch_hybrid_sparse_coo(self):
        """Conventional torch hybrid COO view (sparse_utils.py:436-449).  Kept for API parity; the
        attention path itself never calls it (no coalesce sorts on the hot path)."""
        ret = torch.sparse_coo_tensor(self.indices, self.values, size=tuple(self.shape)).coalesce()
        assert ret.sparse_dim() == 2 and ret.dense_dim() == 1
        return ret


def get_n_nonempty(x, dim, keepdim=False):
    """Number of observations along a sparse dimension (sparse_utils.py:33-43) for a SparseMat:
    read off the CSR/CSC pointers instead of summing a sparse mask."""
    idx = index_for(x)
    ptr = idx.col_ptr if dim == 0 else idx.row_ptr
    ret = (ptr[1:] - ptr[:-1]).to(torch.int64)
    return ret.unsqueeze(dim) if keepdim else ret


def sparse_mean(x, dim, keepdim=False):
    """Mean over the observed entries (sparse_utils.py:91-131) for a SparseMat; rows / columns
    without observations give 0 (what the reference's ``.to_dense()`` of the sparse mean yields).
    dim: 0, 1 or (0, 1)."""
    idx = index_for(x)
    dims = tuple(dim) if isinstance(dim, (tuple, list)) else (dim,)
    if set(dims) == {0, 1}:
        tot = ops.seg_pool(x.values, idx.by_view, idx.row_idx).sum(dim=0) / max(1, x.values.shape[0])
        return tot[None, None, :] if keepdim else tot
    if dims == (0,):
        out = ops.seg_pool(x.values, idx.by_track, idx.col_idx, 1.0, True)
    elif dims == (1,):
        out = ops.seg_pool(x.values, idx.by_view, idx.row_idx, 1.0, True)
    else:
        raise ValueError(dim)
    return out.unsqueeze(dims[0]) if keepdim else out


def sparsemat_feature_cat(sparsemats_list):
    """Concatenate SparseMats of one sparsity pattern along the feature axis (sparse_utils.py:284-307).
    The reference's per-call ``torch.all(indices == ...)`` host syncs become a cheap identity check."""
    assert len(sparsemats_list) > 0
    first = sparsemats_list[0]
    for x in sparsemats_list[1:]:
        assert x.indices is first.indices or x.indices.shape == first.indices.shape
        assert tuple(x.shape[:2]) == tuple(first.shape[:2])
    return first.with_values(torch.cat([x.values for x in sparsemats_list], dim=1))
