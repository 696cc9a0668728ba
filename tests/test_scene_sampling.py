"""sparse-first view sub-sampling (``SceneData.sample_data``, code/datasets/SceneData.py:306-355) against
fixtures generated from the unmodified reference (tests/golden/sample_data.npz): same views drawn for the same numpy
seed, bit-exact indices / counts / graphs, re-normalised values, including the reference's quirk that a NON-consecutive
draw keeps the measurement rows sorted but takes ``Ns`` in draw order."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gasfm_b200.scene import Scene, sample_data
from gasfm_b200.utils import dataset_utils
from oracle import gasfm_cpu

CASES = ["consecutive4", "fraction", "random5"]


def _check(sub, g, name):
    assert np.array_equal(sub.x.indices.cpu().numpy(), g[f"{name}.indices"])
    assert tuple(sub.x.shape) == tuple(g[f"{name}.shape"])
    assert np.array_equal(sub.x.cam_per_pts.cpu().numpy(), g[f"{name}.cam_per_pts"])
    assert np.array_equal(sub.x.pts_per_cam.cpu().numpy(), g[f"{name}.pts_per_cam"])
    assert np.array_equal(sub.y[:, 0, 0].cpu().numpy().astype(np.int64), g[f"{name}.view_ids"])
    assert np.array_equal(sub.Ns.cpu().numpy(), g[f"{name}.Ns"])
    assert np.abs(sub.x.values.cpu().numpy() - g[f"{name}.values"]).max() < 1e-6
    for key, w in sub.graph_wrappers.items():
        assert np.array_equal(w.edge_index.cpu().numpy(), g[f"{name}.graph.{key}"]), key


@pytest.mark.parametrize("name", CASES)
def test_sample_data_matches_reference_golden_sparse_host_scene(name):
    g = load_golden("sample_data")
    M, Ns, y = torch.from_numpy(g["M"]), torch.from_numpy(g["Ns"]), torch.from_numpy(g["y"])
    norm, raw = gasfm_cpu.observation_index(M, Ns), gasfm_cpu.observation_index(M)
    m, n, _ = norm["shape"]
    scene = Scene.from_observations(norm["indices"], norm["values"], m, n, Ns=Ns, obs=raw["values"], y=y)
    num_views, consecutive, seed = g[f"{name}.args"]
    np.random.seed(int(seed))
    sub = sample_data(scene, int(num_views) if num_views >= 1 else float(num_views), consecutive_views=bool(consecutive))
    _check(sub, g, name)
    assert dataset_utils.is_valid_sample(sub) == (sub.x.pts_per_cam.min().item() >= 8)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_sample_data_matches_reference_golden_on_device(name):
    """The same through ``Scene.from_measurements`` (M2sparse kernels) with the scene resident on the GPU."""
    g = load_golden("sample_data")
    dev = "cuda:0"
    scene = Scene.from_measurements(torch.from_numpy(g["M"]).to(dev), torch.from_numpy(g["Ns"]).to(dev), y=torch.from_numpy(g["y"]).to(dev))
    num_views, consecutive, seed = g[f"{name}.args"]
    np.random.seed(int(seed))
    sub = sample_data(scene, int(num_views) if num_views >= 1 else float(num_views), consecutive_views=bool(consecutive))
    assert sub.x.values.is_cuda
    _check(sub, g, name)


def test_subset_views_needs_raw_points_to_reorder():
    idx, vals = gasfm_cpu.synthetic_observations(10, 120, 500, seed=0)
    scene = Scene.from_observations(idx, vals, 10, 120)
    sub = scene.subset_views([2, 3, 4, 5])                 # sorted ids: the normalised values carry over
    assert sub.x.shape[0] == 4 and sub.x.cam_per_pts.min().item() >= 2
    with pytest.raises(ValueError):
        scene.subset_views([5, 2, 3])


@pytest.mark.gpu
def test_core_errors_metric_matches_reference_golden():
    """evaluation.compute_core_errors -> 'our_repro' on the device, against the reference's own value."""
    from gasfm_b200.config import ConfigTree
    from gasfm_b200.evaluation import compute_core_errors
    g = load_golden("core_errors")
    dev = "cuda:0"
    scene = Scene.from_measurements(torch.from_numpy(g["M"]).to(dev), torch.from_numpy(g["Ns"]).to(dev))
    conf = ConfigTree.from_dict({"model": {"view_head": {"enabled": True}, "scenepoint_head": {"enabled": True}}})
    pred = {"Ps_norm": torch.from_numpy(g["Ps_norm"]).to(dev), "pts3D": torch.from_numpy(g["pts3D"]).to(dev)}
    got = compute_core_errors(scene, pred, conf)["our_repro"]
    assert got.is_cuda and got.dim() == 0                   # stays on the device until somebody logs it
    want = float(g["our_repro.f64"])
    assert abs(float(got) - want) < 1e-5 * want

    class DenseData:                                        # duck-typed reference SceneData: dense M, no .obs
        pass
    d = DenseData()
    d.x, d.M, d.Ns_invT = scene.x, torch.from_numpy(g["M"]).to(dev), scene.Ns_invT
    assert abs(float(compute_core_errors(d, pred, conf)["our_repro"]) - want) < 1e-5 * want


def _host_scene(g, y_key="y"):
    M, Ns, y = torch.from_numpy(g["M"]), torch.from_numpy(g["Ns"]), torch.from_numpy(g[y_key])
    norm, raw = gasfm_cpu.observation_index(M, Ns), gasfm_cpu.observation_index(M)
    m, n, _ = norm["shape"]
    return Scene.from_observations(norm["indices"], norm["values"], m, n, Ns=Ns, obs=raw["values"], y=y)


def test_get_subset_matches_reference_golden():
    """SceneData.get_subset (code/datasets/SceneData.py:529-583): the greedy view choice on the observation list."""
    from gasfm_b200.scene import get_subset
    g = load_golden("sample_data")
    sub, chosen = get_subset(_host_scene(g), 5)
    assert sorted(chosen) == g["subset5.view_ids"].tolist()
    _check(sub, g, "subset5")


@pytest.mark.parametrize("name", ["aug_both", "aug_inplane", "aug_tilt"])
def test_rotational_homography_aug_matches_reference_golden(name):
    """SceneData.apply_rotational_homography_aug (code/datasets/SceneData.py:358-453) on the observed points only: same random
    rotations for the same torch seed, cameras and (re-normalised) image points as the reference's dense computation."""
    from gasfm_b200.scene import apply_rotational_homography_aug
    g = load_golden("sample_data")
    scene = _host_scene(g, "aug.y_in")
    inplane, tilt, seed = g[f"{name}.args"]
    torch.manual_seed(int(seed))
    aug = apply_rotational_homography_aug(scene, None if inplane < 0 else float(inplane), None if tilt < 0 else float(tilt))
    assert np.array_equal(aug.x.indices.numpy(), g[f"{name}.indices"])
    assert np.abs(aug.y.numpy() - g[f"{name}.y"]).max() < 1e-4 * max(1.0, np.abs(g[f"{name}.y"]).max())
    assert np.abs(aug.x.values.numpy() - g[f"{name}.values"]).max() < 2e-5
    M = g[f"{name}.M"]
    rows, cols = aug.x.indices[0].numpy(), aug.x.indices[1].numpy()
    pix = np.stack((M[2 * rows, cols], M[2 * rows + 1, cols]), axis=1)
    assert np.abs(aug.obs.numpy() - pix).max() < 1e-2          # pixel coordinates (~1e3) in fp32
