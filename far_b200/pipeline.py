"""The FAR-LoFTR per-pair pose pipeline, batched: what PL_LoFTR.test_step does for one pair
(mp3d_loftr/src/lightning/lightning_loftr.py:325-421), run as N independent B=1 evaluations in one pass.

    LoFTR.forward                       backbone (cuDNN) -> coarse transformer -> dual-softmax matching -> fine level
    estimate_pose_batched               weighted normalised 8-point + cheirality (R,t)   [reference: OpenCV RANSAC on CPU
                                        in a python loop with 3 PCIe hops; here on-device, SURVEY.md 8d config 2]
    for i in range(fine_pred_steps):    FAR head: LoFTR regress layers -> EMM bilinear attention -> gated MLP fusion
        forward_rt_prediction           (the prior-guided 2nd RANSAC round between the two invocations is SURVEY.md
        estimate_pose_batched            8f rank 2 "next"; the solver is simply re-run so the per-pair work of the
                                         recipe of record is preserved)
"""
import torch

from .solver import estimate_pose_batched
from .loftr.pose import pose_mean_6d, pose_std_6d, rotation_6d_to_matrix


class FarPosePipeline:
    def __init__(self, model, K0, K1, fine_pred_steps=None, prior_ransac=False, ransac_kwargs=None):
        self.model = model
        self.K0, self.K1 = K0, K1
        self.steps = fine_pred_steps if fine_pred_steps is not None else model.config.get('fine_pred_steps', 1)
        # prior_ransac=True: the solver call between the two head invocations is the prior-guided RANSAC round of the
        # recipe of record (lightning_loftr.py:343-344, metrics.py:100-123) run on the GPU (far_b200/ransac.py) with the
        # first head prediction as the prior, instead of a plain re-run of the weighted 8-point solver.
        self.prior_ransac = prior_ransac
        self.ransac_kwargs = ransac_kwargs or {}

    @torch.no_grad()
    def __call__(self, image0, image1):
        """image0/1: [N,1,H,W] fp32 CUDA in [0,1].  Returns dict: pose [N,3,4] fused (R|t), regressed_rt [N,9],
        loftr_rt [N,3,4], num_matches [N], gating [N,2]."""
        m = self.model
        data = {'image0': image0, 'image1': image1}
        m(data)
        N = image0.shape[0]
        K0 = self.K0[:N] if self.K0.shape[0] >= N else self.K0.expand(N, 3, 3)
        K1 = self.K1[:N] if self.K1.shape[0] >= N else self.K1.expand(N, 3, 3)
        estimate_pose_batched(data, K0, K1)
        if m.config.get('regress_rt'):
            for i in range(self.steps):
                m.forward_rt_prediction(data)
                if i == 0 and self.steps > 1:
                    if self.prior_ransac:
                        from .ransac import prior_ransac_round
                        rr0 = data['regressed_rt']
                        d0 = rr0.device
                        R0 = rotation_6d_to_matrix(rr0[:, 3:] * pose_std_6d[3:].to(d0) + pose_mean_6d[3:].to(d0))
                        t0 = rr0[:, :3] * pose_std_6d[:3].to(d0) + pose_mean_6d[:3].to(d0)
                        prior_ransac_round(data, K0, K1, torch.cat([R0, t0.unsqueeze(-1)], dim=-1), **self.ransac_kwargs)
                    else:
                        estimate_pose_batched(data, K0, K1)
            rr = data['regressed_rt']
            dev = rr.device
            R = rotation_6d_to_matrix(rr[:, 3:] * pose_std_6d[3:].to(dev) + pose_mean_6d[3:].to(dev))
            t = rr[:, :3] * pose_std_6d[:3].to(dev) + pose_mean_6d[:3].to(dev)
            pose = torch.cat([R, t.unsqueeze(-1)], dim=-1)
        else:
            pose = data['loftr_rt']
        return {'pose': pose, 'regressed_rt': data.get('regressed_rt'), 'loftr_rt': data['loftr_rt'],
                'num_matches': data['num_correspondences_before_ransac'],
                'gating': data.get('gating_reg_weights'), 'data': data}


def shard_pairs(n_pairs, world_size, rank):
    """Contiguous split of a pair batch over ranks: returns (start, stop)."""
    base, rem = divmod(n_pairs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_poses(pose, counts, group=None):
    """The path's only collective: all-gather of [pairs_per_rank, 12] poses + [pairs_per_rank] match counts
    (replaces the reference's pickled gloo gather, mp3d_loftr/src/utils/comm.py:179-219).  Ragged shards are padded
    to the largest shard."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n = torch.tensor([pose.shape[0]], device=pose.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(ns, n, group=group)
    nmax = int(max(x.item() for x in ns))
    buf = torch.zeros((nmax, 13), device=pose.device, dtype=torch.float32)
    buf[:pose.shape[0], :12] = pose.reshape(-1, 12)
    buf[:pose.shape[0], 12] = counts.float()
    out = [torch.zeros_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf, group=group)
    full = torch.cat([o[:int(k.item())] for o, k in zip(out, ns)], 0)
    return full[:, :12].reshape(-1, 3, 4), full[:, 12].round().to(torch.int64)
