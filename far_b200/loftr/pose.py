"""Pose-space helpers used by the FAR head at inference (mp3d_loftr/src/losses/loftr_loss.py:7-39).
Tiny host-side tensor glue ([N,9]-sized); the heavy lifting is in the CUDA ops."""
import torch
import torch.nn.functional as F

# mp3d normalisation constants, loftr_loss.py:7-8
pose_mean_6d = torch.tensor([-0.34898765, 0.17085525, -0.87944315, 0.50275223, 0.03533648, -0.18179045,
                             -0.03533648, 0.98189617, 0.09313615])
pose_std_6d = torch.tensor([1.94014405, 0.36770130, 1.88317520, 0.51837117, 0.12717603, 0.65426397,
                            0.12717603, 0.0188729, 0.09709263])


def rotation_6d_to_matrix(d6):
    """6D -> rotation matrix by Gram-Schmidt, rows (b1, b2, b1 x b2)  (loftr_loss.py:10-29)."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def matrix_to_rotation_6d(m):
    return m[..., :2, :].clone().reshape(*m.size()[:-2], 6)


def compute_normalized_6d(pose_mtx, mean=None, std=None):
    """[t | R[0,:] | R[1,:]] normalised by mean/std (loftr_loss.py:31-39)."""
    mean = pose_mean_6d if mean is None else mean
    std = pose_std_6d if std is None else std
    r6 = matrix_to_rotation_6d(pose_mtx[..., :3, :3])
    v = torch.cat([pose_mtx[..., :3, 3], r6], dim=-1)
    return (v - mean.to(v.device)) / std.to(v.device)
