"""Mirror of the reference's ``ESFMLoss`` (``code/loss_functions.py:69-123``) evaluated sparsely on the
device: only the E observed (view, point) pairs are projected -- the reference builds dense
``[m,3,n]`` tensors (3.6 GB at 1000 x 300k) and synchronises with ``.item()`` inside its gradient hook.
Same constructor keys (``loss.*``), same ``forward(pred_dict, data)`` signature, same value and the same
gradients (including the gradient-equalisation hook) -- see tests/test_gpu_loss.py."""
import torch
from torch import nn

from . import _lib, ops
from .index import index_for


class _EsfmLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Ps, pts3D, obs, index, margin, hinge, hinge_weight, grad_mode):
        ops._require_cuda(Ps, pts3D, obs)
        Ps, pts3D, obs = Ps.contiguous(), pts3D.contiguous(), obs.contiguous()
        E, n, dev = obs.shape[0], pts3D.shape[1], obs.device
        out = torch.empty(2, dtype=torch.float32, device=dev)
        ws = torch.empty(max(1, _lib.size_query("gasfm_esfm_loss_ws_bytes", E) // 4), dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            _lib.call("gasfm_esfm_loss_fwd", _lib.ptr(Ps), _lib.ptr(pts3D), n, _lib.ptr(obs), _lib.ptr(index.row_idx),
                      _lib.ptr(index.col_idx), E, float(margin), int(hinge), float(hinge_weight), _lib.ptr(out), _lib.ptr(ws),
                      _lib.stream_ptr())
        ctx.save_for_backward(Ps, pts3D, obs, out)
        ctx.index, ctx.cfg = index, (float(margin), int(hinge), float(hinge_weight), int(grad_mode))
        return out[0]

    @staticmethod
    def backward(ctx, d_loss):
        Ps, pts3D, obs, stats = ctx.saved_tensors
        index = ctx.index
        margin, hinge, hinge_weight, grad_mode = ctx.cfg
        E, n, dev = obs.shape[0], pts3D.shape[1], obs.device
        G = torch.empty((E, 16), dtype=torch.float32, device=dev)
        up = d_loss.reshape(1).to(torch.float32).contiguous()
        with _lib.device_guard(dev):
            _lib.call("gasfm_esfm_loss_bwd", _lib.ptr(Ps), _lib.ptr(pts3D), n, _lib.ptr(obs), _lib.ptr(index.row_idx),
                      _lib.ptr(index.col_idx), E, margin, hinge, hinge_weight, _lib.ptr(up), _lib.ptr(stats), grad_mode,
                      _lib.ptr(G), _lib.stream_ptr())
        dPs = ops.seg_sum_raw(G[:, :12], index.by_view).view(Ps.shape)
        dpts = ops.seg_sum_raw(G[:, 12:], index.by_track).t().contiguous()
        return dPs, dpts, None, None, None, None, None, None


class ESFMLoss(nn.Module):
    def __init__(self, conf):
        super().__init__()
        assert conf.get_bool('model.view_head.enabled', default=False)
        assert conf.get_bool('model.scenepoint_head.enabled', default=False)
        self.infinity_pts_margin = conf.get_float("loss.infinity_pts_margin")
        self.pts_grad_equalization_pre_perspective_divide = conf.get_bool("loss.pts_grad_equalization_pre_perspective_divide")
        self.normalize_grad_wrt_valid_projections_only = False
        if self.pts_grad_equalization_pre_perspective_divide:
            self.normalize_grad_wrt_valid_projections_only = conf.get_bool("loss.normalize_grad_wrt_valid_projections_only")
        self.hinge_loss = conf.get_bool("loss.hinge_loss")
        self.hinge_loss_weight = conf.get_float("loss.hinge_loss_weight") if self.hinge_loss else 0

    def forward(self, pred_dict, data, epoch=None):
        grad_mode = 0
        if self.pts_grad_equalization_pre_perspective_divide:
            grad_mode = 1 if self.normalize_grad_wrt_valid_projections_only else 2
        x = data.x
        idx = getattr(data, "_gasfm_b200_index", None) or index_for(x)
        return _EsfmLoss.apply(pred_dict["Ps_norm"], pred_dict["pts3D"], x.values, idx, self.infinity_pts_margin,
                               self.hinge_loss, self.hinge_loss_weight, grad_mode)


def get_loss_func(conf):
    """``loss_functions.get_loss_func`` for the losses provided here (code/loss_functions.py:8-21)."""
    spec = conf.get_string('loss.func')
    if spec != 'ESFMLoss':
        raise NotImplementedError("gasfm_b200 provides ESFMLoss; %s stays with the reference implementation" % spec)
    return ESFMLoss(conf)
