"""Prediction cache between the LoFTR stage (mp3d_loftr) and the 8pt-ViT head (interiornetStreetlearn_8ptVit)
-- SURVEY.md 8(f) rank 4.  Same on-disk format and coordinate convention as the reference, plus an in-memory path so
the ViT head can consume solver poses produced on the GPU in the same run instead of one `torch.load` per pair.

Reference format (mp3d_loftr/src/lightning/lightning_loftr.py:348-360):
    <parent>/<split>/loftr_preds/<pair_id>.pt                  torch tensor [3,4] float64  = [R | t]
    <parent>/<split>/loftr_num_correspondences/<pair_id>.pt    0-d torch tensor (number of correspondences)
Reader (interiornetStreetlearn_8ptVit/test_streetlearn_interiornet.py:250-267,
        interiornetStreetlearn_8ptVit/src/data_readers/interiornet.py:117-126):
    T = eye(4); T[:3] = preds;  T = F1 T F1^-1 (mp3d axis flip);  T = F2 T F2^-1 (interiornet axis swap);  float64
    missing files -> preds = eye(4)[:3], num_corr = 0.
`ViTEss.forward(..., loftr_preds=[B,4,4] (or [B,3,4]), loftr_num_corr=[B])` takes the result as is.
"""
import os

import torch

# T -> F T F^-1 for F1 = diag(1,-1,-1,1) (mp3d) then F2 = [[0,1,0,0],[1,0,0,0],[0,0,-1,0],[0,0,0,1]] (interiornet)
_F1 = torch.tensor([[1., 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=torch.float64)
_F2 = torch.tensor([[0., 1, 0, 0], [1, 0, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=torch.float64)


def _paths(parent, split, pair_id):
    return (os.path.join(parent, split, 'loftr_preds', f'{pair_id}.pt'),
            os.path.join(parent, split, 'loftr_num_correspondences', f'{pair_id}.pt'))


def save_prediction(parent, split, pair_id, R, t, num_corr=None):
    """Writes one pair exactly as PL_LoFTR.test_step does (lightning_loftr.py:348-360): preds = cat([R, t[:,None]], 1)
    as float64 [3,4]; num_corr as a 0-d tensor (skipped when None, like NO_SAVE_NUMCORR)."""
    pp, np_ = _paths(parent, split, pair_id)
    os.makedirs(os.path.dirname(pp), exist_ok=True)
    os.makedirs(os.path.dirname(np_), exist_ok=True)
    R = torch.as_tensor(R).detach().cpu().to(torch.float64).reshape(3, 3)
    t = torch.as_tensor(t).detach().cpu().to(torch.float64).reshape(3, 1)
    torch.save(torch.cat([R, t], dim=1), pp)
    if num_corr is not None:
        torch.save(torch.as_tensor(int(num_corr)), np_)


def save_batch(parent, split, pair_ids, poses, num_corr):
    """poses [N,3,4] (any device / dtype), num_corr [N]: one file pair per id (the reference's layout)."""
    poses = torch.as_tensor(poses).detach().cpu().to(torch.float64)
    num_corr = torch.as_tensor(num_corr).detach().cpu()
    for k, pid in enumerate(pair_ids):
        save_prediction(parent, split, pid, poses[k, :, :3], poses[k, :, 3], int(num_corr[k]))


def to_vit_convention(poses):
    """[N,3,4] or [N,4,4] solver poses in the mp3d_loftr camera convention -> [N,4,4] float64 in the convention the
    8pt-ViT head was trained on (the two conjugations of test_streetlearn_interiornet.py:258-264).  Runs on the
    tensor's device: no host round trip for poses that were produced on the GPU."""
    poses = torch.as_tensor(poses)
    dev = poses.device
    n = poses.shape[0]
    T = torch.eye(4, dtype=torch.float64, device=dev).repeat(n, 1, 1)
    T[:, :3, :] = poses[:, :3, :].to(torch.float64)
    F1, F2 = _F1.to(dev), _F2.to(dev)
    T = F1 @ T @ torch.linalg.inv(F1)
    T = F2 @ T @ torch.linalg.inv(F2)
    return T


def load_prediction(parent, split, pair_id, device='cpu'):
    """One pair as the reference's test script reads it (test_streetlearn_interiornet.py:250-267): returns
    (loftr_preds [1,4,4] float64 in the ViT convention, loftr_num_corr [1]); identity [1,3,4] float32 / 0 when either
    file is missing."""
    pp, np_ = _paths(parent, split, pair_id)
    if os.path.exists(pp) and os.path.exists(np_):
        preds = torch.load(pp).unsqueeze(0)
        nc = torch.load(np_).unsqueeze(0)
        return to_vit_convention(preds).to(device), nc.to(device)
    return torch.eye(4)[:3].unsqueeze(0).to(device), torch.tensor([0]).to(device)


def load_batch(parent, split, pair_ids, device='cpu'):
    """Batched reader for ViTEss.forward: (loftr_preds [N,4,4] float64, loftr_num_corr [N] int64).  Missing pairs get
    the identity pose and 0 correspondences, as in the reference."""
    Ts, ncs = [], []
    for pid in pair_ids:
        T, nc = load_prediction(parent, split, pid)
        if T.shape[-2] == 3:
            T = torch.cat([T.to(torch.float64), torch.tensor([[[0., 0, 0, 1]]], dtype=torch.float64)], dim=1)
        Ts.append(T[0])
        ncs.append(nc.reshape(-1)[0].to(torch.int64))
    return torch.stack(Ts).to(device), torch.stack(ncs).to(device)


def from_pipeline(out):
    """FarPosePipeline output -> (loftr_preds [N,4,4] float64, loftr_num_corr [N]) for ViTEss.forward, on the device
    the poses live on (BASELINE configs[2]: "8pt-ViT + cached-correspondence solver" fed from the same run).  The
    count is `num_correspondences_after_ransac`, the quantity the reference's cache writer stores
    (lightning_loftr.py:356-359), so the in-memory path and save_batch -> load_batch give the same tensors."""
    return to_vit_convention(out['loftr_rt']), out['num_inliers'].to(torch.int64)


def save_pipeline(parent, split, pair_ids, out):
    """Write a FarPosePipeline output in the reference's on-disk format (poses = the solver's loftr_rt, counts =
    num_correspondences_after_ransac; lightning_loftr.py:348-360)."""
    save_batch(parent, split, pair_ids, out['loftr_rt'], out['num_inliers'])
