"""FAR fusion head of the map-free model (mapfree_6dreg/lib/models/regression/model.py:198-233 `regression_mlp`,
with the MLP definitions of :66-84): same parameter names (`pose_regressor.{0,2,4}`, `moe_predictor.{0,2,4}`) so a
`RegressionModel` checkpoint's head weights load, GEMMs on the sm_100a kernels.

Scope note: the rest of RegressionModel (ResUNet encoder, correlation-volume aggregator,
nn.TransformerEncoder) is the 8(f) rank-1/3 "next" work and stays on cuDNN/cuBLAS in the reference harness; its
LoFTR matcher is `far_b200.loftr.LoFTR(upstream_loftr_cfg())` and its solver `far_b200.solver`."""
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_RELU, ACT_SIGMOID


def compute_6d(r):
    """lib/utils/loss.py:9."""
    return r[..., :2, :].clone().reshape(*r.size()[:-2], 6)


class RegressionHead(nn.Module):
    def __init__(self, use_prior=True):
        super().__init__()
        self.H2, self.H, self.pose_size = 512, 256 * 12 * 9, 9
        self.use_prior = use_prior
        self.num_corr_size = 3
        self.pose_regressor = nn.Sequential(nn.Linear(self.H, self.H2), nn.ReLU(), nn.Linear(self.H2, self.H2),
                                            nn.ReLU(), nn.Linear(self.H2, self.pose_size))
        self.moe_predictor = nn.Sequential(nn.Linear(self.H + 2 * self.pose_size + self.num_corr_size, self.H2),
                                           nn.ReLU(), nn.Linear(self.H2, self.H2), nn.ReLU(), nn.Linear(self.H2, 2),
                                           nn.Sigmoid())

    def regression_mlp(self, features, loftr_preds, loftr_num_corr, R=None, t=None):
        """features [B,256,12,9]; loftr_preds [B,3,4]; loftr_num_corr [B,3] (use_prior) or [B,1]/[B].
        Returns (R6d [B,6], t [B,3]) like the reference."""
        B = features.shape[0]
        dev = features.device
        lp = loftr_preds.float().to(dev)
        l9 = torch.cat([lp[..., 3], compute_6d(lp[..., :3, :3])], dim=-1)
        nc = loftr_num_corr.detach().float().to(dev) / 500
        if not self.use_prior:
            if nc.dim() == 1:
                nc = nc.unsqueeze(0)
            nc = torch.cat([nc, nc / 10, nc / 100], dim=-1)  # vanilla-RANSAC filler (:209-212)
        feats = features.reshape(B, -1)
        pr, mp = self.pose_regressor, self.moe_predictor
        pred = ops.linear(ops.linear(ops.linear(feats, pr[0].weight, pr[0].bias, ACT_RELU), pr[2].weight, pr[2].bias,
                                     ACT_RELU), pr[4].weight, pr[4].bias)
        ratio = torch.linalg.norm(pred[..., :3], dim=-1) / torch.clamp(torch.linalg.norm(l9[..., :3], dim=-1), 1e-2, 1e2)
        lt = l9[..., :3] * torch.clamp(ratio.unsqueeze(1), 1e-2, 1e2)
        lout = torch.cat([lt, l9[..., 3:], nc], dim=-1)
        hid = ops.linear_cat_tail(feats, torch.cat([pred, lout], dim=-1), mp[0], ACT_RELU)
        wt = ops.linear(ops.linear(hid, mp[2].weight, mp[2].bias, ACT_RELU), mp[4].weight, mp[4].bias, ACT_SIGMOID)
        t_out = wt[..., :1] * pred[..., :3] + (1 - wt[..., :1]) * lout[..., :3]
        R_out = wt[..., 1:] * pred[..., 3:] + (1 - wt[..., 1:]) * lout[..., 3:-self.num_corr_size]
        return R_out, t_out
