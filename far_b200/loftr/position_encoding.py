"""2-D sinusoidal position encoding (mp3d_loftr/src/loftr/utils/position_encoding.py:6-42), incl. the
`temp_bug_fix=False` variant where `-log(1e4) / d_model // 2` floor-divides (:28).  The table is module state
(a non-persistent buffer, as in the reference); the add + NCHW->NLC flatten is one CUDA kernel."""
import math

import torch
from torch import nn

from .. import ops


class PositionEncodingSine(nn.Module):
    def __init__(self, d_model, max_shape=(256, 256), temp_bug_fix=True):
        super().__init__()
        pe = torch.zeros((d_model, *max_shape))
        y_position = torch.ones(max_shape).cumsum(0).float().unsqueeze(0)
        x_position = torch.ones(max_shape).cumsum(1).float().unsqueeze(0)
        k = torch.arange(0, d_model // 2, 2).float()
        if temp_bug_fix:
            div_term = torch.exp(k * (-math.log(10000.0) / (d_model // 2)))
        else:
            div_term = torch.exp(k * (-math.log(10000.0) / d_model // 2))
        div_term = div_term[:, None, None]
        pe[0::4] = torch.sin(x_position * div_term)
        pe[1::4] = torch.cos(x_position * div_term)
        pe[2::4] = torch.sin(y_position * div_term)
        pe[3::4] = torch.cos(y_position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0), persistent=False)  # [1, C, H, W]
        self._hwc = {}

    def table_hwc(self, h, w):
        key = (h, w, self.pe.device)
        if key not in self._hwc:
            self._hwc[key] = self.pe[0, :, :h, :w].permute(1, 2, 0).reshape(h * w, -1).contiguous()
        return self._hwc[key]

    def forward_flatten(self, x):
        """x [N,C,H,W] -> (x + pe) rearranged to [N, H*W, C] (position_encoding.py:37-42 + loftr.py:100-101)."""
        return ops.pos_encode_flatten(x, self.table_hwc(x.size(2), x.size(3)))

    def forward(self, x):
        n, c, h, w = x.shape
        return self.forward_flatten(x).reshape(n, h, w, c).permute(0, 3, 1, 2)
