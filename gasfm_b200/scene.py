"""Scene container holding what the model reads from the reference's ``SceneData``
(``code/datasets/SceneData.py``): ``x`` (SparseMat of normalised observations) and
``graph_wrappers`` (the four aggregation graphs).  Sparse-first: a scene can be built straight
from an observation list without ever forming the dense ``M[2m,n]``."""
import copy

import torch

from .utils import dataset_utils
from .utils.sparse_utils import SparseMat


class Scene:
    def __init__(self, x, scene_name="scene", y=None):
        self.x = x
        self.scene_name = scene_name
        self.y = y
        self.device = x.values.device
        self.graph_wrappers = dataset_utils.create_axial_aggregation_graphs(x)

    @classmethod
    def from_measurements(cls, M, Ns, scene_name="scene", y=None):
        """Dense path, as ``SceneData.__init__`` does it: ``x = M2sparse(M, normalize=True, Ns)``
        (SceneData.py:43)."""
        return cls(dataset_utils.M2sparse(M, normalize=True, Ns=Ns), scene_name, y)

    @classmethod
    def from_observations(cls, indices, values, m, n, scene_name="scene"):
        """Sparse path: ``indices [2,E]`` int64 row-major sorted, ``values [E,2]`` fp32."""
        indices = torch.as_tensor(indices, dtype=torch.int64)
        values = torch.as_tensor(values, dtype=torch.float32)
        cam_per_pts = torch.bincount(indices[1], minlength=n).unsqueeze(1)
        pts_per_cam = torch.bincount(indices[0], minlength=m).unsqueeze(1)
        return cls(SparseMat(values, indices, cam_per_pts, pts_per_cam, (m, n, 2)), scene_name)

    def to(self, device, *args, **kwargs):
        kwargs.pop("dense_on_demand", None)
        ret = copy.copy(self)
        ret.x = self.x.to(device, **kwargs)
        ret.graph_wrappers = {k: w.to(device, **kwargs) for k, w in self.graph_wrappers.items()}
        if torch.is_tensor(self.y):
            ret.y = self.y.to(device, **kwargs)
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
