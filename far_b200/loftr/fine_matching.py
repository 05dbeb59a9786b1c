"""FineMatching (mp3d_loftr/src/loftr/utils/fine_matching.py:8-76): one warp per match."""
import math

import torch
import torch.nn as nn

from .. import ops


class FineMatching(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config

    def forward(self, feat_f0, feat_f1, data, train=False):
        M, WW, C = feat_f0.shape
        W = int(math.sqrt(WW))
        scale = data['hw0_i'][0] / data['hw0_f'][0]
        self.M, self.W, self.WW, self.C, self.scale = M, W, WW, C, scale
        if M == 0:  # no coarse matches (:33-41)
            data.update({'expec_f': torch.empty(0, 3, device=feat_f0.device),
                         'mkpts0_f': data['mkpts0_c'], 'mkpts1_f': data['mkpts1_c']})
            return
        if 'scale0' in data:
            raise NotImplementedError("per-image scale1 is outside the FAR eval path")
        expec_f, mkpts1_f = ops.fine_match(feat_f0, feat_f1, data['mkpts1_c'], (W // 2) * scale)
        data.update({'expec_f': expec_f, 'mkpts0_f': data['mkpts0_c'], 'mkpts1_f': mkpts1_f})
