"""Map-free submission writer (second half): the text format mapfree_6dreg/submission.py:30-82
produces -- one `pose_<scene>.txt` per scene inside a zip, one line per query frame

    <query image name> qw qx qy qz tx ty tz inliers

floats with 6 decimals (`Pose.__str__`, :37-42), the rotation as a unit quaternion with w >= 0 (transforms3d `mat2quat`,
:67), `inliers` = data['inliers'][.,0] printed as the reference prints it (an int from the solver, `0.0` for the
no-pose fallback, model.py:257-262), frames whose pose is NaN / inf skipped (:59-61), lines joined by '\n' without a
trailing newline (:78-79).  Takes the batched device tensors RegressionModel.forward returns (R in the 6-D
representation, t [B,3]): a whole batch is converted on the device and leaves it once."""
import zipfile
from collections import defaultdict

import torch

from .mapfree import rotation_6d_to_matrix


def matrix_to_quaternion_wxyz(R):
    """[B,3,3] rotation matrices -> [B,4] unit quaternions (w, x, y, z), w >= 0 (transforms3d.quaternions.mat2quat's
    convention, the one submission.py:45 uses).  Branch-free: picks the best-conditioned of the four candidate forms."""
    m = R.double()
    m00, m01, m02 = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    m10, m11, m12 = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    m20, m21, m22 = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    q = torch.stack([
        torch.stack([1 + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, 1 + m00 - m11 - m22, m01 + m10, m02 + m20], -1),
        torch.stack([m02 - m20, m01 + m10, 1 - m00 + m11 - m22, m12 + m21], -1),
        torch.stack([m10 - m01, m02 + m20, m12 + m21, 1 - m00 - m11 + m22], -1)], 1)      # [B,4 candidates,4]
    best = torch.stack([q[:, 0, 0], q[:, 1, 1], q[:, 2, 2], q[:, 3, 3]], 1).argmax(1)
    qq = q[torch.arange(q.shape[0], device=q.device), best]
    qq = qq / qq.norm(dim=1, keepdim=True)
    return torch.where(qq[:, :1] < 0, -qq, qq)


def poses_to_lines(frame_paths, R6d, t, inliers):
    """One submission line per pair (submission.py:37-42); (index, line) for the frames with a finite pose."""
    R = rotation_6d_to_matrix(R6d.detach().double())
    q = matrix_to_quaternion_wxyz(R).cpu()
    t = t.detach().double().reshape(-1, 3).cpu()
    c = torch.as_tensor(inliers).detach().double().cpu()
    c = c[:, 0] if c.dim() == 2 else c.reshape(-1)
    out = []
    for i, fp in enumerate(frame_paths):
        if not (torch.isfinite(q[i]).all() and torch.isfinite(t[i]).all()):
            continue
        v = float(c[i])
        conf = str(int(v)) if v != 0 and v == int(v) else str(v)
        nums = " ".join(f"{float(x):.6f}" for x in list(q[i]) + list(t[i]))
        out.append((i, f"{fp} {nums} {conf}"))
    return out


class SubmissionWriter:
    """Accumulates batches and writes the zip the map-free benchmark scorer reads (one pose_<scene>.txt per scene,
    save_submission :75-79)."""

    def __init__(self):
        self.by_scene = defaultdict(list)

    def add_batch(self, scene_ids, frame_paths, R6d, t, inliers):
        for i, line in poses_to_lines(frame_paths, R6d, t, inliers):
            self.by_scene[scene_ids[i]].append(line)

    def write(self, path):
        with zipfile.ZipFile(path, "w") as z:
            for scene, lines in self.by_scene.items():
                z.writestr(f"pose_{scene}.txt", "\n".join(lines).encode("utf-8"))
        return path
