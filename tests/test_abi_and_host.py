"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the host
mirror has the reference's module tree / state_dict names, and the CUDA-only product path refuses
to run on CPU tensors (no silent fallback)."""
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import MODEL_FIXTURES, ROOT, golden_model, load_golden
from gasfm_b200 import _lib
from gasfm_b200.config import ConfigTree, gasfm_conf
from gasfm_b200.models.graph_attn_sfm import GraphAttnSfMNet
from gasfm_b200.scene import Scene
from gasfm_b200.utils import dataset_utils
from oracle import gasfm_cpu


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gasfm_b200.h")).read()
    return sorted(set(re.findall(r"GASFM_API[^;(]*?\b(gasfm_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _declared_symbols()
    assert len(syms) >= 20
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gasfm_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes prototypes out of sync with the header"
    assert lib.gasfm_abi_version() == 1


def test_tensor_core_kernel_shape_predicates():
    """Host-only shape predicates that steer the dispatch in gasfm_b200.ops (no kernel is launched): the scaled 2xFP16
    GEMM takes 64 <= K <= 256, everything else stays on 3xTF32; the fp16 weight gradient takes Nout in {128, 256} and
    Kout a multiple of 64; row strides must keep rows 16-byte aligned."""
    lib = _lib.load()
    E = 495592
    assert lib.gasfm_linear_f16x2_supported(E, 256, 256, 256, 256) == 1
    assert lib.gasfm_linear_f16x2_supported(E, 256, 64, 64, 768) == 1           # grouped output, ldc = 3 N
    assert lib.gasfm_linear_f16x2_supported(E, 32, 32, 32, 32) == 0             # shipped d = 32: narrow K
    assert lib.gasfm_linear_f16x2_supported(E, 256, 512, 512, 256) == 0         # K > 256: tile does not fit the registers
    assert lib.gasfm_linear_f16x2_supported(E, 40, 256, 256, 40) == 0           # N % 16
    assert lib.gasfm_linear_tf32x3_supported(E, 32, 32, 32, 32) == 1
    assert lib.gasfm_linear_tf32x3_supported(E, 256, 768, 256, 256) == 1        # concatenated dY (K = 3 d)
    assert lib.gasfm_linear_tf32x3_supported(E, 256, 2, 2, 256) == 0            # first block (K = 2): write-bound kernel instead
    assert lib.gasfm_linear_tf32x3_supported(E, 256, 256, 255, 256) == 0        # misaligned rows
    assert lib.gasfm_wgrad_f16x2_supported(E, 256, 256, 256, 256) == 1
    assert lib.gasfm_wgrad_f16x2_supported(E, 256, 64, 256, 64) == 1
    assert lib.gasfm_wgrad_f16x2_supported(E, 64, 64, 64, 64) == 0              # -> wgrad_small / 3xTF32
    assert lib.gasfm_wgrad_f16x2_supported(E, 256, 32, 256, 32) == 0
    assert lib.gasfm_wgrad_tf32x3_supported(E, 256, 32, 256, 32) == 1
    assert lib.gasfm_wgrad_small_supported(64, 64, 64, 64) == 1
    # LayerNorm inside the GEMM with the normalised operand as a by-product: only the CTA-pair kernel (N = K = 256) writes it
    assert lib.gasfm_linear_f16x2_ln_y_supported(E, 256, 256, 256, 768) == 1
    assert lib.gasfm_linear_f16x2_ln_y_supported(E, 128, 128, 128, 384) == 0
    assert lib.gasfm_linear_f16x2_ln_y_supported(E, 256, 192, 192, 768) == 0
    assert lib.gasfm_update_bwd_views_ws_bytes(100, 256) > 0
    # workspace queries are pure host arithmetic too
    assert lib.gasfm_wgrad_f16x2_ws_bytes(256, 256) == (148 * 256 * 256 + 148 * 256) * 4
    assert lib.gasfm_col_sum_ws_bytes(100, 256) == 0 and lib.gasfm_col_sum_ws_bytes(50000, 256) == 195 * 256 * 4


def test_ln_relu_width_predicate_mirrors_the_kernel():
    """Widths the fused LayerNorm+ReLU kernel rejects (> 256 and not a multiple of 4; csrc/edge_ops.cu) must take the
    torch fallback in the layers instead of raising."""
    from gasfm_b200 import ops
    assert all(ops.ln_relu_width_supported(w) for w in (1, 2, 3, 32, 64, 255, 256, 260, 512, 1024))
    assert not any(ops.ln_relu_width_supported(w) for w in (0, 257, 258, 514, 1023, 1028, 2048))


def test_header_cites_the_reference_for_each_entry_point():
    text = open(os.path.join(ROOT, "include", "gasfm_b200.h")).read()
    for cite in ("utils/dataset_utils.py:86-113", "utils/sparse_utils.py:436-449", "models/layers.py:329-335",
                 "utils/sparse_utils.py:406-419", "models/layers.py:941-945"):
        assert cite in text


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_state_dict_names_and_shapes_match_reference(name):
    conf, params, _ = golden_model(name)
    model = GraphAttnSfMNet(ConfigTree.from_dict(conf))
    sd = model.state_dict()
    assert sorted(sd) == sorted(params)
    for k, v in params.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    model.load_state_dict(params, strict=True)


def test_shipped_conf_parameter_count():
    model = GraphAttnSfMNet(gasfm_conf())
    assert len(model.state_dict()) == 886                              # SURVEY.md section 6 [probe]
    assert sum(p.numel() for p in model.parameters()) == 145_165_560


def test_product_path_refuses_cpu_tensors():
    model = GraphAttnSfMNet(gasfm_conf(n_feat_view=32, n_feat_global=64, num_layers=2))
    idx, vals = gasfm_cpu.synthetic_observations(10, 60, 300, seed=0)
    scene = Scene.from_observations(idx, vals, 10, 60)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(scene)


def test_graph_wrappers_match_reference_golden():
    g = load_golden("index_build")
    idx = torch.from_numpy(g["indices"])
    m, n = int(g["shape"][0]), int(g["shape"][1])
    scene = Scene.from_observations(idx, g["values"], m, n)
    assert np.array_equal(scene.x.cam_per_pts.numpy(), g["cam_per_pts"])
    assert np.array_equal(scene.x.pts_per_cam.numpy(), g["pts_per_cam"])
    for name, w in scene.graph_wrappers.items():
        assert np.array_equal(w.edge_index.numpy(), g[f"{name}.edge_index"]), name
        assert np.array_equal(w.valid_indices.numpy(), g[f"{name}.valid_indices"]), name
        assert [w.m, w.n, w.agg_dim, w.n_agg_nodes] == list(g[f"{name}.meta"]), name
    w = scene.graph_wrappers["proj2view"]
    feats = torch.arange(idx.shape[1] * 3, dtype=torch.float32).reshape(-1, 3)
    nodes = w.generate_node_features(scene.x.with_values(feats))
    assert nodes.shape == (idx.shape[1] + m, 3) and torch.equal(nodes[: idx.shape[1]], feats)
    assert w.extract_target_node_features(nodes).shape == (m, 1, 3)


def test_plugin_registers_reference_module_names():
    from gasfm_b200 import plugin
    saved = {k: sys.modules.get(k) for k in ("models", "models.graph_attn_sfm", "models.layers", "models.baseNet",
                                             "models.SetOfSet")}
    try:
        plugin.install()
        mod = __import__("models.graph_attn_sfm")
        cls = getattr(getattr(mod, "graph_attn_sfm"), "GraphAttnSfMNet")   # what general_utils.get_class does
        assert cls is GraphAttnSfMNet
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_m2sparse_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="CUDA"):
        dataset_utils.M2sparse(torch.zeros(4, 5))


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="live reference only exists in the build container")
def test_module_tree_equals_live_reference_shipped_conf():
    from oracle import ref_import
    ref = ref_import.import_reference()
    conf = gasfm_conf()
    theirs = ref.graph_attn_sfm.GraphAttnSfMNet(ref.ConfigTree.from_dict(conf))
    ours = GraphAttnSfMNet(conf)
    a, b = theirs.state_dict(), ours.state_dict()
    assert list(a) == list(b)                                          # same names, same ORDER
    assert all(a[k].shape == b[k].shape for k in a)
    ours.load_state_dict(a, strict=True)


def test_activation_recompute_plan():
    """Which blocks recompute their [E, d] activations (pure host arithmetic): nothing while everything fits in half of the
    device, otherwise the LAST blocks keep theirs as far as the budget allows; BASELINE.json configs[2] at d = 256 on a 180 GB
    part as measured (1 GPU: 11 of 12 recompute; half the tracks per GPU: 3 of 12; a quarter: none)."""
    from gasfm_b200 import ops
    total = 191_500_000_000
    E = 4_987_789
    plan = lambda e, **kw: ops.recompute_plan(e, 256, 12, total, **{"mode": "auto", "keep": "auto", **kw})   # noqa: E731
    assert plan(E) == 11 and plan(E // 2) == 3 and plan(E // 4) == 0 and plan(E // 8) == 0
    assert plan(E, grad_enabled=False) == 0                       # inference keeps nothing anyway
    assert plan(495_592) == 0                                     # cfg2
    assert plan(E, mode="off") == 0 and plan(495_592, mode="on") == 12
    assert plan(E, keep="3") == 9 and plan(495_592, mode="on", keep="1") == 11 and plan(E, keep="99") == 0
    assert plan(3 * E) == 12                                      # nothing fits: every block recomputes
