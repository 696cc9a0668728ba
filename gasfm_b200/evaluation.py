"""Mirror of the reference's per-step metric ``evaluation.compute_core_errors`` (``code/evaluation.py:8-74``, called every
training step at ``code/train.py:91``) without host round trips: the reference moves the predictions to numpy and
recomputes a dense ``[m, n]`` reprojection on the CPU; here one kernel walks the E observed pairs on the device and
the result stays a device scalar (``float()`` it when it is actually logged)."""
import torch

from . import _lib
from .index import index_for


def reprojection_error(Ps, pts3D, obs, index):
    """nan-mean over the observations of ``|| obs_e - pflat(Ps[i] @ pts3D[:, j])[:2] ||``
    (geo_utils.reprojection_error_with_points, utils/geo_utils.py:371-391) -> (mean, count) device scalars."""
    Ps, pts3D, obs = Ps.detach().contiguous().float(), pts3D.detach().contiguous().float(), obs.contiguous().float()
    if not Ps.is_cuda:
        raise RuntimeError("gasfm_b200.evaluation runs on CUDA tensors only (no CPU fallback)")
    E, n, dev = obs.shape[0], pts3D.shape[1], obs.device
    out = torch.empty(2, dtype=torch.float32, device=dev)
    ws = torch.empty(max(1, _lib.size_query("gasfm_esfm_loss_ws_bytes", E) // 4), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.call("gasfm_reproj_error", _lib.ptr(Ps), _lib.ptr(pts3D), n, _lib.ptr(obs), _lib.ptr(index.row_idx),
                  _lib.ptr(index.col_idx), E, _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr())
    return out[0], out[1]


def _raw_observations(data):
    """[E,2] raw image points in the order of ``data.x``: ``Scene.obs``, or gathered from a dense ``data.M [2m,n]``."""
    obs = getattr(data, "obs", None)
    if obs is not None:
        return obs
    M = data.M
    m = M.shape[0] // 2
    rows, cols = data.x.indices[0], data.x.indices[1]
    return M.reshape(m, 2, M.shape[1])[rows, :, cols]


def compute_core_errors(data, pred_dict, conf):
    """-> {'our_repro': 0-d device tensor}: mean reprojection error of the predicted cameras and points in image
    coordinates (evaluation.py:27-32).  The depth-head metric (``eval.calc_reprojerr_with_gtposes_for_depth_pred``,
    evaluation.py:34-72) is outside the attention path and stays with the reference implementation."""
    core_errors = {}
    view_head = conf.get_bool('model.view_head.enabled', default=False)
    scenepoint_head = conf.get_bool('model.scenepoint_head.enabled', default=False)
    if conf.get_bool('eval.calc_reprojerr_with_gtposes_for_depth_pred', default=False):
        raise NotImplementedError("gasfm_b200.evaluation covers 'our_repro'; the depth-head metric stays with the reference")
    if view_head and scenepoint_head:
        Ns_inv = data.Ns_invT.transpose(1, 2).to(pred_dict['Ps_norm'].device)
        Ps = Ns_inv @ pred_dict['Ps_norm'].detach()                      # unnormalised cameras
        X = pred_dict['pts3D'].detach()
        X = X / X[-1, :]                                                 # geo_utils.pflat
        core_errors['our_repro'], _ = reprojection_error(Ps, X, _raw_observations(data), index_for(data.x))
    return core_errors
