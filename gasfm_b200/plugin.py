"""Register gasfm_b200's mirror modules under the names the reference imports.

The reference instantiates its model through
``general_utils.get_class("models." + conf.get_string("model.type"))(conf)``
(``code/main.py:134-136``, ``code/utils/general_utils.py:84-90``) with
``model.type = "graph_attn_sfm.GraphAttnSfMNet"``.  After ``install()``, ``import models.graph_attn_sfm``
(and ``models.layers``, ``models.baseNet``, ``models.SetOfSet``) resolve to this package, so the
reference's drivers pick up the B200 implementation without source changes.  See INTEGRATION.md."""
import importlib
import sys
import types

_MAP = {
    "models.graph_attn_sfm": "gasfm_b200.models.graph_attn_sfm",
    "models.layers": "gasfm_b200.models.layers",
    "models.baseNet": "gasfm_b200.models.baseNet",
    "models.SetOfSet": "gasfm_b200.models.SetOfSet",
}


def install():
    """Make ``models.*`` resolve to gasfm_b200.  The reference's own data containers (``utils.sparse_utils.SparseMat``,
    ``utils.dataset_utils`` graph wrappers, ``datasets.SceneData``) stay in place: the models duck-type them
    (``tests/test_gpu_model.py::test_accepts_duck_typed_reference_containers``) and build their CSR/CSC index on first use."""
    mapping = dict(_MAP)
    pkg = sys.modules.get("models")
    if pkg is None:
        pkg = types.ModuleType("models")
        pkg.__path__ = []
        sys.modules["models"] = pkg
    for alias, target in mapping.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], leaf, mod)
    return mapping
