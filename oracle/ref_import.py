"""Import the UNMODIFIED reference sources from ``/root/reference/code`` (oracle, test-only).

Only usable in the build container: ``/root/reference`` does not exist on the GPU
box, so nothing that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may
call into this module.  Its one consumer is ``tests/golden/make_golden.py``, which
turns reference outputs into committed fixtures, plus a container-only test that
re-validates ``oracle/gasfm_cpu.py`` against the live reference when it is present.

The reference imports several packages that are absent from this image.  They are
replaced by the smallest stand-ins that let the hot-path modules import:

* ``torch_geometric.nn.GATv2Conv``  -> ``oracle.gatv2conv.GATv2Conv`` (restated algorithm)
* ``pytorch3d.transforms``          -> quaternion / 6d / axis-angle conversions (restated
  from the documented pytorch3d formulas; used at ``code/models/baseNet.py:43,48``)
* ``pyhocon``                       -> a dict-backed ``ConfigTree`` with ``get_*(key, default=)``
* ``cvxpy``, ``dask``, ``plotly``, ``matplotlib``: import-time only, empty modules
* ``datasets`` / ``models`` / ``utils``: namespace shims pointing into the reference tree
  (the installed HuggingFace ``datasets`` package would otherwise shadow ``code/datasets``)
"""
import importlib
import os
import sys
import types

import torch

REF_CODE = "/root/reference/code"


def reference_available():
    return os.path.isdir(REF_CODE)


# ----------------------------------------------------------------------------------------
# pytorch3d.transforms stand-ins
# ----------------------------------------------------------------------------------------
def quaternion_to_matrix(quaternions):
    """Real-first quaternion -> rotation matrix, scale 2/|q|^2 (pytorch3d semantics)."""
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = torch.nn.functional.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def axis_angle_to_matrix(axis_angle):
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = angles * 0.5
    small = angles.abs() < 1e-6
    sin_half_over_angle = torch.where(small, 0.5 - angles * angles / 48, torch.sin(half) / torch.where(small, torch.ones_like(angles), angles))
    quats = torch.cat([torch.cos(half), axis_angle * sin_half_over_angle], dim=-1)
    return quaternion_to_matrix(quats)


# ----------------------------------------------------------------------------------------
# pyhocon stand-in
# ----------------------------------------------------------------------------------------
_MISSING = object()


class ConfigTree(dict):
    """Nested dict with pyhocon's dotted ``get_*(key, default=)`` accessors."""

    def _lookup(self, key, default=_MISSING):
        node = self
        for part in key.split("."):
            if isinstance(node, dict) and part in node:
                node = node[part]
            else:
                if default is _MISSING:
                    raise KeyError(key)
                return default
        return node

    def get(self, key, default=_MISSING):
        return self._lookup(key, default)

    def get_int(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else int(v)

    def get_float(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else float(v)

    def get_bool(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else bool(v)

    def get_string(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else str(v)

    def get_list(self, key, default=_MISSING):
        return self._lookup(key, default)

    def get_config(self, key, default=_MISSING):
        return self._lookup(key, default)

    @classmethod
    def from_dict(cls, d):
        out = cls()
        for k, v in d.items():
            out[k] = cls.from_dict(v) if isinstance(v, dict) else v
        return out


class _ConfigFactory:
    @staticmethod
    def from_dict(d):
        return ConfigTree.from_dict(d)

    @staticmethod
    def parse_file(path):
        raise NotImplementedError("HOCON parsing is not part of the oracle; build configs with from_dict")

    parse_string = parse_file


def _empty_module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave like a package so ``import a.b`` works
    return mod


_installed = False


def install_stubs():
    """Register the stand-in modules; idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"{REF_CODE} is not present; the live reference can only be imported in the build container")
    from oracle.gatv2conv import GATv2Conv

    tg = _empty_module("torch_geometric")
    tg_nn = _empty_module("torch_geometric.nn", GATv2Conv=GATv2Conv)
    tg.nn = tg_nn
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.nn"] = tg_nn

    p3d = _empty_module("pytorch3d")
    p3d_tr = _empty_module(
        "pytorch3d.transforms",
        quaternion_to_matrix=quaternion_to_matrix,
        rotation_6d_to_matrix=rotation_6d_to_matrix,
        axis_angle_to_matrix=axis_angle_to_matrix,
    )
    p3d.transforms = p3d_tr
    sys.modules["pytorch3d"] = p3d
    sys.modules["pytorch3d.transforms"] = p3d_tr

    class _HOCONConverter:  # import-time only
        pass

    sys.modules["pyhocon"] = _empty_module(
        "pyhocon", ConfigFactory=_ConfigFactory, ConfigTree=ConfigTree, HOCONConverter=_HOCONConverter
    )

    for name in ("cvxpy", "dask", "dask.array", "plotly", "plotly.express", "plotly.subplots",
                 "plotly.graph_objects", "matplotlib", "matplotlib.image"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _empty_module(name)
    if "plotly.subplots" in sys.modules and not hasattr(sys.modules["plotly.subplots"], "make_subplots"):
        sys.modules["plotly.subplots"].make_subplots = lambda *a, **k: None
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        tb = _empty_module("torch.utils.tensorboard", SummaryWriter=object)
        sys.modules["torch.utils.tensorboard"] = tb

    # Namespace shims: the reference's code/{models,utils,datasets} have no __init__.py.
    for pkg in ("models", "utils", "datasets"):
        mod = types.ModuleType(pkg)
        mod.__path__ = [os.path.join(REF_CODE, pkg)]
        sys.modules[pkg] = mod
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    _installed = True


def import_reference():
    """Return a namespace with the reference modules the hot path touches."""
    install_stubs()
    ns = types.SimpleNamespace()
    ns.sparse_utils = importlib.import_module("utils.sparse_utils")
    ns.dataset_utils = importlib.import_module("utils.dataset_utils")
    ns.geo_utils = importlib.import_module("utils.geo_utils")
    ns.SceneData = importlib.import_module("datasets.SceneData")
    ns.layers = importlib.import_module("models.layers")
    ns.graph_attn_sfm = importlib.import_module("models.graph_attn_sfm")
    ns.SetOfSet = importlib.import_module("models.SetOfSet")
    ns.ConfigTree = ConfigTree
    return ns
