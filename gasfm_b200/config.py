"""Minimal stand-in for pyhocon's ``ConfigTree`` (pyhocon is not required): a nested dict with the
dotted ``get_int / get_bool / get_string / get_float(key, default=)`` accessors the models use
(reference: ``code/models/graph_attn_sfm.py:12-41``, ``code/models/baseNet.py:12-14``).  A real
pyhocon ``ConfigTree`` can be passed to the models just as well."""

_MISSING = object()


class ConfigTree(dict):
    def _lookup(self, key, default=_MISSING):
        node = self
        for part in key.split("."):
            if isinstance(node, dict) and part in node:
                node = node[part]
            elif default is _MISSING:
                raise KeyError(key)
            else:
                return default
        return node

    def get(self, key, default=_MISSING):
        return self._lookup(key, default)

    def _typed(self, key, default, cast):
        v = self._lookup(key, default)
        return v if v is None else cast(v)

    def get_int(self, key, default=_MISSING):
        return self._typed(key, default, int)

    def get_float(self, key, default=_MISSING):
        return self._typed(key, default, float)

    def get_bool(self, key, default=_MISSING):
        return self._typed(key, default, bool)

    def get_string(self, key, default=_MISSING):
        return self._typed(key, default, str)

    def get_list(self, key, default=_MISSING):
        return self._lookup(key, default)

    @classmethod
    def from_dict(cls, d):
        out = cls()
        for k, v in d.items():
            out[k] = cls.from_dict(v) if isinstance(v, dict) else v
        return out


def gasfm_conf(**model_overrides):
    """The shipped Euclidean GASFM model configuration
    (``code/confs/gasfm/learning_euc_noaug_gasfm.conf:43-75``) with optional overrides of the
    ``model`` section; ``calibrated=False`` switches ``dataset.calibrated``."""
    calibrated = model_overrides.pop("calibrated", True)
    model = dict(type="graph_attn_sfm.GraphAttnSfMNet", n_heads=4, stateful_global_features=True,
                 global2view_and_global2scenepoint_enabled=False, n_feat_proj=32, n_feat_scenepoint=64,
                 n_feat_view=1024, n_feat_global=2048, num_layers=12,
                 n_hidden_layers_scenepoint_update=0, n_hidden_layers_view_update=0,
                 n_hidden_layers_global_update=0, n_hidden_layers_proj_update=0,
                 use_norm_proj_update=True, add_residual_skipconn_proj_update=True,
                 add_skipconn_from_init_projfeat=True, pos_emb_n_freq=0,
                 depth_head=dict(enabled=False, n_feat=128, n_hidden_layers=2),
                 view_head=dict(enabled=True, n_hidden_layers=2, rot_representation="quat"),
                 scenepoint_head=dict(enabled=True, n_hidden_layers=2))
    model.update(model_overrides)
    return ConfigTree.from_dict(dict(dataset=dict(calibrated=calibrated), model=model))
