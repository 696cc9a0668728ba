import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_FIXTURES = sorted(f[len("model_"):-len(".npz")] for f in os.listdir(GOLDEN) if f.startswith("model_"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_model(name):
    """-> (conf dict, params {key: tensor f32}, fixture npz)"""
    g = load_golden("model_" + name)
    conf = json.loads(str(g["conf_json"]))
    params = {k[len("param."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")}
    return conf, params, g


def grad_errors(got, want):
    """Worst per-parameter gradient error, each scaled by max(|g_param|, 1e-3 * largest gradient):
    analytically-zero gradients (softmax shift invariance) are rounding noise only."""
    gscale = 1e-3 * max(float(np.abs(v).max()) for v in want.values())
    worst, worst_key = 0.0, None
    for k, w in want.items():
        err = float(np.abs(got[k] - w).max()) / max(gscale, float(np.abs(w).max()))
        if err > worst:
            worst, worst_key = err, k
    return worst, worst_key
