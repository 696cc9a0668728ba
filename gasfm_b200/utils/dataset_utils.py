"""Mirror of the reference's ``code/utils/dataset_utils.py`` pieces that feed the attention path:
``get_M_valid_points`` (:86-113), ``M2sparse`` (:116-156) and ``AxialAggregationGraphWrapper``
(:464-597).  The index work runs in CUDA kernels (``gasfm_m2sparse_*``); inputs that live on the
host are moved to the GPU for it and the results are returned on the input's device."""
import copy

import torch

from .. import _lib
from .constants import MIN_N_VIEWS_PER_POINT, MIN_N_POINTS_PER_VIEW  # noqa: F401
from .sparse_utils import SparseMat


def sample_indices(N, num_samples, adjacent):
    """View ids of a training sample (dataset_utils.py:25-40): all views, a fraction / number of consecutive views, or
    a random subset -- same numpy RNG calls as the reference, so equal seeds give equal samples."""
    import numpy as np

    if num_samples == 1:
        return np.arange(N)
    if num_samples < 1:
        num_samples = int(np.ceil(num_samples * N))
    num_samples = max(2, num_samples)
    if num_samples >= N:
        return np.arange(N)
    if adjacent:
        start = np.random.randint(0, N - num_samples + 1)
        return np.arange(start, start + num_samples)
    return np.random.choice(N, num_samples, replace=False)


def is_valid_sample(data):
    """dataset_utils.py:12-13."""
    return (data.x.pts_per_cam.min().item() >= MIN_N_POINTS_PER_VIEW and
            data.x.cam_per_pts.min().item() >= MIN_N_VIEWS_PER_POINT)


def _cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError("gasfm_b200: building the observation index needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _valid_and_counts(M):
    """-> (M on GPU [2m,n] f32, valid u8 [m,n], cam_per_pts i64 [n], pts_per_cam i64 [m], E)"""
    assert M.dim() == 2 and M.shape[0] % 2 == 0
    dev = M.device if M.is_cuda else _cuda_device()
    Mg = M.to(device=dev, dtype=torch.float32).contiguous()
    m, n = Mg.shape[0] // 2, Mg.shape[1]
    valid = torch.empty((m, n), dtype=torch.uint8, device=dev)
    cam_per_pts = torch.empty(n, dtype=torch.int64, device=dev)
    pts_per_cam = torch.empty(m, dtype=torch.int64, device=dev)
    n_obs = torch.empty(1, dtype=torch.int64, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_m2sparse_count", _lib.ptr(Mg), m, n, MIN_N_VIEWS_PER_POINT, _lib.ptr(valid),
                  _lib.ptr(cam_per_pts), _lib.ptr(pts_per_cam), _lib.ptr(n_obs), _lib.stream_ptr())
    return Mg, valid, cam_per_pts, pts_per_cam, int(n_obs.item())


def get_M_valid_points(M):
    """[2m,n] (or [m,n,2]) measurements -> [m,n] bool validity mask (dataset_utils.py:86-113)."""
    if M.dim() == 3:
        M = M.permute(0, 2, 1).reshape(M.shape[0] * 2, M.shape[1])
    _, valid, _, _, _ = _valid_and_counts(M)
    return valid.to(torch.bool).to(M.device)


def M2sparse(M, normalize=False, Ns=None):
    """Dense [2m,n] measurement matrix -> SparseMat [m,n,2] (dataset_utils.py:116-156).
    With ``normalize`` the values are (Ns @ [x;y;1])[:2] (geo_utils.normalize_M, geo_utils.py:689-703)."""
    src_dev = M.device
    Mg, valid, cam_per_pts, pts_per_cam, E = _valid_and_counts(M)
    dev = Mg.device
    m, n = Mg.shape[0] // 2, Mg.shape[1]
    indices = torch.empty((2, E), dtype=torch.int64, device=dev)
    values = torch.empty((E, 2), dtype=torch.float32, device=dev)
    Ng = None
    if normalize:
        assert Ns is not None
        Ng = Ns.to(device=dev, dtype=torch.float32).contiguous()
    ws = torch.empty(max(1, _lib.size_query("gasfm_m2sparse_ws_bytes", m, n) // 8), dtype=torch.int64, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_m2sparse_fill", _lib.ptr(Mg), _lib.ptr(Ng), _lib.ptr(valid), m, n, E, _lib.ptr(indices),
                  _lib.ptr(values), _lib.ptr(ws), _lib.stream_ptr())
    out = SparseMat(values, indices, cam_per_pts.unsqueeze(1), pts_per_cam.unsqueeze(1), (m, n, 2))
    return out if src_dev == dev else out.to(src_dev)


class AxialAggregationGraphWrapper:
    """Row-wise / column-wise aggregation graph over the valid matrix elements
    (dataset_utils.py:464-597).  Same attributes and methods as the reference; ``edge_index`` is
    materialised lazily because the CUDA path works from CSR/CSC segments and never reads it."""

    def __init__(self, m, n, agg_dim, valid_indices=None, device=None):
        if device is not None and valid_indices is not None:
            assert valid_indices.device == torch.device(device)
        if valid_indices is not None:
            self.device = valid_indices.device
        elif device is not None:
            self.device = torch.device(device)
        else:
            raise AssertionError("either valid_indices or device must be given")
        self.m, self.n = m, n
        assert agg_dim in (0, 1)
        self.agg_dim = agg_dim
        self.non_agg_dim = 1 - agg_dim
        self.n_agg_nodes = (m, n)[self.non_agg_dim]
        self.dense = valid_indices is None
        if self.dense:
            rows = torch.arange(m, dtype=torch.int64, device=self.device)
            cols = torch.arange(n, dtype=torch.int64, device=self.device)
            valid_indices = torch.cartesian_prod(rows, cols).T.contiguous()
        assert valid_indices.dtype == torch.int64 and valid_indices.dim() == 2 and valid_indices.shape[0] == 2
        self.valid_indices = valid_indices
        self._edge_index = None

    @property
    def edge_index(self):
        """[2,E] int64: source = element id, target = E + row/col id (dataset_utils.py:511-537)."""
        if self._edge_index is None or self._edge_index.device != self.valid_indices.device:
            self._edge_index = self.create_sparse_axial_aggregation_edges()
        return self._edge_index

    def create_sparse_axial_aggregation_edges(self):
        n_el = self.valid_indices.shape[1]
        src = torch.arange(n_el, dtype=torch.int64, device=self.valid_indices.device)
        return torch.stack((src, n_el + self.valid_indices[self.non_agg_dim]))

    def generate_node_features(self, M, x_agg=None):
        """cat(element features [E,d], aggregation-node features [T,d] or zeros) (dataset_utils.py:539-576).
        API parity only: the CUDA path never concatenates."""
        if isinstance(M, SparseMat):
            x_el = M.values
        elif M.is_sparse:
            x_el = M.values()
        else:
            x_el = M.reshape(self.m * self.n, -1)
        if x_agg is None:
            x_agg = torch.zeros((self.n_agg_nodes, x_el.shape[1]), dtype=torch.float32, device=x_el.device)
        return torch.cat((x_el, x_agg), dim=0)

    def extract_target_node_features(self, x):
        """Last T rows, shaped (1,n,d) or (m,1,d) (dataset_utils.py:578-590)."""
        if self.agg_dim == 0:
            return x[None, -self.n:, :]
        return x[-self.m:, None, :]

    def to(self, device, **kwargs):
        ret = copy.copy(self)
        ret.device = device
        ret.valid_indices = ret.valid_indices.to(device, **kwargs)
        ret._edge_index = None if self._edge_index is None else self._edge_index.to(device, **kwargs)
        return ret


def create_axial_aggregation_graphs(x):
    """The four graphs of SceneData.create_axial_aggregation_graphs (code/datasets/SceneData.py:153-239)
    for a SparseMat ``x``: proj2view, proj2scenepoint, view2global (views with >= 8 points),
    scenepoint2global (tracks with >= 2 views)."""
    m, n = x.shape[0], x.shape[1]
    idx = x.indices
    dev = idx.device
    pts_per_view = x.pts_per_cam.reshape(-1)
    views_per_pt = x.cam_per_pts.reshape(-1)
    view_rows = torch.nonzero(pts_per_view >= MIN_N_POINTS_PER_VIEW)[:, 0]
    pt_cols = torch.nonzero(views_per_pt >= MIN_N_VIEWS_PER_POINT)[:, 0]
    v2g = torch.stack((view_rows, torch.zeros_like(view_rows)))
    p2g = torch.stack((torch.zeros_like(pt_cols), pt_cols))
    return {
        "proj2view": AxialAggregationGraphWrapper(m, n, 1, valid_indices=idx),
        "proj2scenepoint": AxialAggregationGraphWrapper(m, n, 0, valid_indices=idx),
        "view2global": AxialAggregationGraphWrapper(m, 1, 0, valid_indices=v2g, device=dev),
        "scenepoint2global": AxialAggregationGraphWrapper(1, n, 1, valid_indices=p2g, device=dev),
    }
