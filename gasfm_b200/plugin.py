"""Register gasfm_b200's mirror modules under the names the reference imports.

The reference instantiates its model through
``general_utils.get_class("models." + conf.get_string("model.type"))(conf)``
(``code/main.py:134-136``, ``code/utils/general_utils.py:84-90``) with
``model.type = "graph_attn_sfm.GraphAttnSfMNet"``.  After ``install()``, ``import models.graph_attn_sfm``
(and ``models.layers``, ``models.baseNet``, ``utils.sparse_utils``) resolve to this package, so the
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


def install(override_utils=False):
    """Make ``models.*`` resolve to gasfm_b200.  With ``override_utils`` the sparse containers
    (``utils.sparse_utils.SparseMat``, ``utils.dataset_utils.M2sparse`` / graph wrapper) are swapped
    too; otherwise the reference's own containers are used and duck-typed by the model."""
    mapping = dict(_MAP)
    if override_utils:
        mapping["utils.sparse_utils"] = "gasfm_b200.utils.sparse_utils"
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
