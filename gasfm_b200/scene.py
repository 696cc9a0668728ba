"""Scene container holding what the model, the loss and the per-step metric read from the reference's ``SceneData``
(``code/datasets/SceneData.py``): ``x`` (SparseMat of normalised observations), ``graph_wrappers`` (the four
aggregation graphs), and optionally the raw image points of the same observations, ``Ns`` and ``y``.  Sparse-first: a
scene can be built straight from an observation list without ever forming the dense ``M[2m,n]``, and view sub-sampling
(``sample_data``, SceneData.py:306-355) works on the observation list."""
import copy

import numpy as np
import torch

from .utils import dataset_utils
from .utils.constants import MIN_N_VIEWS_PER_POINT
from .utils.sparse_utils import SparseMat


class Scene:
    def __init__(self, x, scene_name="scene", y=None, Ns=None, obs=None, calibrated=True):
        self.x = x
        self.scene_name = scene_name
        self.y = y                      # [m,3,4] ground-truth cameras (SceneData.y) or None
        self.Ns = Ns                    # [m,3,3] normalisation matrices (SceneData.Ns) or None
        self.obs = obs                  # [E,2] raw image points of the observations, in x's order, or None
        self.calibrated = calibrated
        self.device = x.values.device
        self.graph_wrappers = dataset_utils.create_axial_aggregation_graphs(x)

    @property
    def Ns_invT(self):
        """transpose(inverse(Ns)) as SceneData stores it (SceneData.py:49)."""
        return torch.transpose(torch.inverse(self.Ns), 1, 2)

    @classmethod
    def from_measurements(cls, M, Ns, scene_name="scene", y=None, calibrated=True):
        """Dense path, as ``SceneData.__init__`` does it: ``x = M2sparse(M, normalize=True, Ns)`` (SceneData.py:43);
        the raw image points are kept sparsely (same observations, same order) instead of the dense ``M``."""
        x = dataset_utils.M2sparse(M, normalize=True, Ns=Ns)
        raw = dataset_utils.M2sparse(M, normalize=False)
        return cls(x, scene_name, y, Ns, raw.values, calibrated)

    @classmethod
    def from_observations(cls, indices, values, m, n, scene_name="scene", Ns=None, obs=None, y=None, calibrated=True):
        """Sparse path: ``indices [2,E]`` int64 row-major sorted, ``values [E,2]`` fp32 normalised image points
        (``obs``: the raw ones, needed by the reprojection metric and by re-normalising view subsets)."""
        indices = torch.as_tensor(indices, dtype=torch.int64)
        values = torch.as_tensor(values, dtype=torch.float32)
        cam_per_pts = torch.bincount(indices[1], minlength=n).unsqueeze(1)
        pts_per_cam = torch.bincount(indices[0], minlength=m).unsqueeze(1)
        obs = None if obs is None else torch.as_tensor(obs, dtype=torch.float32)
        return cls(SparseMat(values, indices, cam_per_pts, pts_per_cam, (m, n, 2)), scene_name, y, Ns, obs, calibrated)

    def to(self, device, *args, **kwargs):
        """``SceneData.to`` (SceneData.py:241-264).  ``dense_on_demand`` is accepted and meaningless: there are no
        dense members to leave behind."""
        kwargs.pop("dense_on_demand", None)
        ret = copy.copy(self)
        ret.x = self.x.to(device, **kwargs)
        ret.graph_wrappers = {}
        for k, w in self.graph_wrappers.items():
            if w.valid_indices is self.x.indices:
                # proj2view / proj2scenepoint list the observations themselves: share the moved tensor, do not copy it again
                w = copy.copy(w)
                w.device, w.valid_indices, w._edge_index = device, ret.x.indices, None
                ret.graph_wrappers[k] = w
            else:
                ret.graph_wrappers[k] = w.to(device, **kwargs)
        for key in ("y", "Ns", "obs"):
            v = getattr(self, key)
            if torch.is_tensor(v):
                setattr(ret, key, v.to(device, **kwargs))
        ret.device = device
        return ret

    def prepare(self):
        """Build everything the model derives from the scene on first use -- the CSR/CSC observation index, the
        chunk tables and the two single-target global plans -- now, so that the first forward issues no host
        synchronisation (needed before CUDA-graph capture and by ranks that share one GPU)."""
        from .index import index_for
        from .models.layers import plan_for

        idx = index_for(self.x)
        idx.shard = getattr(self, "shard", None)
        for key in ("view2global", "scenepoint2global"):
            plan_for(self.graph_wrappers[key])
        return self

    def _buffers(self):
        """The tensors a step reads from the scene, by name (aliases of ``x.indices`` listed once)."""
        out = {"values": self.x.values, "indices": self.x.indices, "cam_per_pts": self.x.cam_per_pts, "pts_per_cam": self.x.pts_per_cam}
        for k in ("view2global", "scenepoint2global"):
            out[k] = self.graph_wrappers[k].valid_indices
        for k in ("y", "Ns", "obs"):
            if torch.is_tensor(getattr(self, k)):
                out[k] = getattr(self, k)
        return out

    def signature(self):
        """What has to agree for one scene to be written over another in place (``copy_from``): the shapes of every
        buffer -- i.e. (m, n, E) and the number of valid views / tracks of the two global graphs."""
        return tuple(self.x.shape) + tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(self._buffers().items()))

    def copy_from(self, other, non_blocking=True):
        """Overwrite this scene's tensors IN PLACE with ``other``'s (same ``signature()``; typically a pinned host scene
        into the device scene a captured step reads, ``gasfm_b200.graphs.StreamedStep``).  Cached indices / plans on this
        scene become stale: the captured step rebuilds them at the same addresses."""
        if self.signature() != other.signature():
            raise ValueError("Scene.copy_from: the scenes differ in shape (m, n, E or the number of valid views / tracks)")
        src = other._buffers()
        for k, dst in self._buffers().items():
            dst.copy_(src[k], non_blocking=non_blocking)
        self.scene_name = other.scene_name
        return self

    def invalidate(self):
        """Forget the cached observation index and global plans (after the index tensors were rewritten in place)."""
        from .index import _INDEX_ATTR
        from .models.layers import _PLAN_ATTR
        for holder in (self, self.x):
            if hasattr(holder, _INDEX_ATTR):
                delattr(holder, _INDEX_ATTR)
        for w in self.graph_wrappers.values():
            if hasattr(w, _PLAN_ATTR):
                delattr(w, _PLAN_ATTR)
        return self

    def pin_memory(self):
        """Page-lock the host tensors so that ``.to(device, non_blocking=True)`` is a true async copy."""
        self.x.values = self.x.values.pin_memory()
        self.x.indices = self.x.indices.pin_memory()
        self.x.cam_per_pts = self.x.cam_per_pts.pin_memory()
        self.x.pts_per_cam = self.x.pts_per_cam.pin_memory()
        for w in self.graph_wrappers.values():
            if w.valid_indices is not self.x.indices:
                w.valid_indices = w.valid_indices.pin_memory()
            else:
                w.valid_indices = self.x.indices
        for k in ("proj2view", "proj2scenepoint"):
            self.graph_wrappers[k].valid_indices = self.x.indices
        return self

    # -- view sub-sampling ----------------------------------------------------------------------
    def subset_views(self, view_ids):
        """The scene restricted to ``view_ids``, as ``SceneData.sample_data`` builds it (SceneData.py:306-355) but on
        the observation list: the selected rows of ``M`` in ascending view order, tracks seen by fewer than
        ``MIN_N_VIEWS_PER_POINT`` of the selected views dropped and the rest renumbered, observations re-normalised.
        Like the reference, ``Ns`` / ``y`` are taken in the order of ``view_ids`` as given (``data.Ns[indices]``) while
        the measurement rows are sorted (``M_indices = np.sort(...)``) -- identical for consecutive views, the shipped case."""
        view_ids = np.asarray(view_ids, dtype=np.int64).reshape(-1)
        m, n = self.x.shape[0], self.x.shape[1]
        dev = self.x.indices.device
        order = torch.as_tensor(np.sort(view_ids), device=dev)
        if order.numel() > 1 and bool((order[1:] == order[:-1]).any()):
            raise ValueError("view_ids must be distinct")
        new_row_of = torch.full((m,), -1, dtype=torch.int64, device=dev)
        new_row_of[order] = torch.arange(order.numel(), device=dev)
        rows, cols = self.x.indices[0], self.x.indices[1]
        keep = new_row_of[rows] >= 0
        views_per_track = torch.bincount(cols[keep], minlength=n)
        track_ok = views_per_track >= MIN_N_VIEWS_PER_POINT
        keep &= track_ok[cols]
        new_col_of = torch.cumsum(track_ok.to(torch.int64), 0) - 1
        sel = torch.nonzero(keep)[:, 0]
        new_idx = torch.stack((new_row_of[rows[sel]], new_col_of[cols[sel]]))          # row-major order is preserved
        n_new = int(track_ok.sum().item())
        taken = torch.as_tensor(view_ids, device=dev)
        Ns = None if self.Ns is None else self.Ns[taken]
        y = None if self.y is None else self.y[taken.to(self.y.device)]
        obs = None if self.obs is None else self.obs[sel]
        if obs is not None and Ns is not None:
            hom = torch.cat((obs, torch.ones_like(obs[:, :1])), dim=1)                 # (Ns @ [x; y; 1])[:2], geo_utils.py:689-703
            values = torch.einsum("eij,ej->ei", Ns[new_idx[0]][:, :2, :], hom)
        else:
            if not np.array_equal(view_ids, np.sort(view_ids)):
                raise ValueError("re-ordering views needs the raw image points and Ns (Scene.obs / Scene.Ns)")
            values = self.x.values[sel]
        cam_per_pts = torch.bincount(new_idx[1], minlength=n_new).unsqueeze(1)
        pts_per_cam = torch.bincount(new_idx[0], minlength=order.numel()).unsqueeze(1)
        x = SparseMat(values, new_idx, cam_per_pts, pts_per_cam, (int(order.numel()), n_new, 2))
        return Scene(x, self.scene_name, y, Ns, obs, self.calibrated)


def sample_data(data, num_views, consecutive_views=True):
    """``SceneData.sample_data`` (SceneData.py:306-355): a random subset of the views (numpy RNG, like the reference's
    ``dataset_utils.sample_indices``), built sparsely on the scene's own device."""
    indices = dataset_utils.sample_indices(data.x.shape[0], num_views, adjacent=consecutive_views)
    return data.subset_views(indices)


def get_subset(data, subset_size):
    """``SceneData.get_subset`` (SceneData.py:529-583): greedy choice of ``subset_size`` views -- start with the view that sees the
    most points, then repeatedly add the view sharing the most points with those seen so far (first index on ties, selected
    views count as sharing nothing) -- evaluated on the observation list (one bincount over the edges per added view instead
    of a dense ``[m, n]`` mask product).  -> (Scene restricted to the chosen views, chosen view ids in selection order)."""
    m, n = data.x.shape[0], data.x.shape[1]
    rows, cols = data.x.indices[0], data.x.indices[1]
    dev = rows.device
    available = torch.ones(m, dtype=torch.bool, device=dev)
    first = int(torch.bincount(rows, minlength=m).argmax().item())
    seen = torch.zeros(n, dtype=torch.bool, device=dev)
    seen[cols[rows == first]] = True
    available[first] = False
    chosen = [first]
    for _ in range(subset_size - 1):
        shared = torch.bincount(rows, weights=(seen[cols] & available[rows]).to(torch.float64), minlength=m)
        nxt = int(shared.argmax().item())
        seen[cols[(rows == nxt) & available[rows]]] = True
        available[nxt] = False
        chosen.append(nxt)
    return data.subset_views(np.sort(np.asarray(chosen, dtype=np.int64))), chosen


def apply_rotational_homography_aug(data, inplane_rot_aug_max_angle=None, tilt_rot_aug_max_angle=None):
    """``SceneData.apply_rotational_homography_aug`` (SceneData.py:358-453): a random in-plane and / or tilt rotation per view,
    applied to the cameras (``y <- Ns^-1 R Ns y``) and to the image points (``pflat(Ns^-1 R Ns [x; y; 1])``), on the E observed
    points only (the reference transforms the dense ``[m, 3, n]`` array).  Draws the same ``torch.rand`` numbers in the same
    order as the reference, so equal seeds on the same device give equal augmentations.  Needs ``Scene.obs`` and ``Scene.Ns``."""
    import math

    if inplane_rot_aug_max_angle is None and tilt_rot_aug_max_angle is None:
        return data
    if data.obs is None or data.Ns is None:
        raise ValueError("the rotational augmentation needs the raw image points and Ns (Scene.obs / Scene.Ns)")
    dev = data.obs.device
    m, n = data.x.shape[0], data.x.shape[1]
    R = torch.eye(3, device=dev)[None].repeat(m, 1, 1)
    inplane = inplane_rot_aug_max_angle or 0
    assert inplane >= 0
    if inplane > 0:
        angle = inplane * (2 * torch.rand((m,), dtype=torch.float32, device=dev) - 1)
        rotvec = torch.zeros((m, 3), dtype=torch.float32, device=dev)
        rotvec[:, 2] = angle / 180. * math.pi
        R = _axis_angle_to_matrix(rotvec) @ R
    tilt = tilt_rot_aug_max_angle or 0
    assert tilt >= 0
    if tilt > 0:
        angle = tilt * (2 * torch.rand((m,), dtype=torch.float32, device=dev) - 1)
        alpha = torch.rand((m,), dtype=torch.float32, device=dev) * 2 * math.pi
        axis = torch.zeros((m, 3), dtype=torch.float32, device=dev)
        axis[:, 0], axis[:, 1] = torch.cos(alpha), torch.sin(alpha)
        R = _axis_angle_to_matrix(axis * angle[:, None] / 180. * math.pi) @ R
    Ns = data.Ns.to(dev)
    Ns_inv = torch.linalg.inv(Ns)
    y = None if data.y is None else (Ns_inv @ R @ Ns) @ data.y.to(dev)
    rows = data.x.indices[0]
    hom = torch.cat((data.obs, torch.ones_like(data.obs[:, :1])), dim=1)                   # old pixel coordinates [E,3]
    norm_old = torch.einsum("eij,ej->ei", Ns[rows], hom)
    norm_new = torch.einsum("eij,ej->ei", R[rows], norm_old)
    pix_new = torch.einsum("eij,ej->ei", Ns_inv[rows], norm_new)
    obs = pix_new[:, :2] / pix_new[:, 2:3]                                                 # geo_utils.batch_pflat
    values = torch.einsum("eij,ej->ei", Ns[rows][:, :2, :], torch.cat((obs, torch.ones_like(obs[:, :1])), dim=1))
    x = SparseMat(values, data.x.indices, data.x.cam_per_pts, data.x.pts_per_cam, tuple(data.x.shape),
                  _index=getattr(data.x, "_gasfm_b200_index", None))
    return Scene(x, data.scene_name, y, data.Ns, obs, data.calibrated)


def _axis_angle_to_matrix(axis_angle):
    """pytorch3d.transforms.axis_angle_to_matrix (Rodrigues via quaternions, as pytorch3d does it)."""
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = angles * 0.5
    small = angles.abs() < 1e-6
    sin_half_over_angle = torch.where(small, 0.5 - (angles * angles) / 48, torch.sin(half) / torch.where(small, torch.ones_like(angles), angles))
    q = torch.cat((torch.cos(half), axis_angle * sin_half_over_angle), dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))
