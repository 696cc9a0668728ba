"""Output decoding shared by the SfM networks; mirror of ``code/models/baseNet.py:8-92``
(tiny per-view math, plain torch)."""
import abc

import torch
from torch.nn import functional as F


def quaternion_to_matrix(q):
    """Real-first quaternion -> rotation, scaled by 2/|q|^2 (what ``pytorch3d.transforms.
    quaternion_to_matrix`` computes for the un-normalised head output, baseNet.py:48)."""
    r, i, j, k = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                     s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                     s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def project_to_rot(m):
    """Nearest rotation by SVD (code/utils/geo_utils.py:25-31)."""
    u, _, v = torch.svd(m)
    vt = v.transpose(1, 2)
    det = torch.det(u @ vt).view(-1, 1, 1)
    return u @ torch.cat((vt[:, :2, :], vt[:, -1:, :] * det), 1)


class BaseNet(torch.nn.Module):
    def __init__(self, conf):
        super().__init__()
        self.calibrated = conf.get_bool('dataset.calibrated')
        self.normalize_output = conf.get_string('model.view_head.normalize_output', default=None)
        self.rot_representation = conf.get_string('model.view_head.rot_representation', default='quat')
        self.soft_sign = torch.nn.Softsign()
        if self.calibrated:
            channels = {'6d': 9, 'quat': 7, 'svd': 12}
            if self.rot_representation not in channels:
                raise ValueError("Illegal output format: %r" % (self.rot_representation,))
            self.out_channels = channels[self.rot_representation]
        else:
            self.out_channels = 12

    @abc.abstractmethod
    def forward(self, data):
        pass

    def extract_depth_outputs(self, depth_out):
        return {"depths": depth_out}

    def extract_view_outputs(self, x):
        if self.calibrated:
            if self.rot_representation == '6d':
                rot = rotation_6d_to_matrix(x[:, :6])
            elif self.rot_representation == 'svd':
                rot = project_to_rot(x[:, :9].reshape(-1, 3, 3))
            else:
                rot = quaternion_to_matrix(x[:, :4])
            Ps = torch.cat((rot, x[:, -3:].unsqueeze(-1)), dim=-1)
        else:
            Ps = x.reshape(-1, 3, 4)
            if self.normalize_output == "Chirality":
                Ps = Ps * (torch.sign(Ps[:, :3, :3].det()) / Ps[:, 2, :3].norm(dim=1)).reshape(-1, 1, 1)
            elif self.normalize_output == "Differentiable Chirality":
                Ps = Ps * (self.soft_sign(Ps[:, :3, :3].det() * 10e3) / Ps[:, 2, :3].norm(dim=1)).reshape(-1, 1, 1)
            elif self.normalize_output == "Frobenius":
                Ps = Ps / Ps.norm(dim=(1, 2), p='fro', keepdim=True)
        return {"Ps_norm": Ps}

    def extract_scenepoint_outputs(self, pts_3D):
        ones = torch.ones(1, pts_3D.shape[1], device=pts_3D.device, dtype=pts_3D.dtype)
        return {"pts3D": torch.cat((pts_3D, ones))}
