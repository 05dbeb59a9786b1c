"""CoarseMatching (mp3d_loftr/src/loftr/utils/coarse_matching.py:58-265), dual-softmax branch, eval.

The [N, L, S] confidence matrix (92 MB/pair at 640x480) is never materialised unless asked for
(`config['materialize_conf_matrix']`, default False): the kernels compute row/column log-sum-exps from 128x128
score tiles, recompute the tiles to select mutual-nearest matches, and compact them in (b, i) order."""
import torch
import torch.nn as nn

from .. import ops
from .._lib import ENGINE_AUTO


class CoarseMatching(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.thr = config['thr']
        self.border_rm = config['border_rm']
        self.train_coarse_percent = config['train_coarse_percent']
        self.train_pad_num_gt_min = config['train_pad_num_gt_min']
        self.match_type = config['match_type']
        if self.match_type != 'dual_softmax':
            # the sinkhorn branch imports a superglue.py that is not in the reference tree (:74-77)
            raise NotImplementedError("only match_type='dual_softmax' is runnable in the reference")
        self.temperature = config['dsmax_temperature']
        self.materialize_conf_matrix = bool(config.get('materialize_conf_matrix', False))
        self.engine = ENGINE_AUTO

    def forward_begin(self, feat_c0, feat_c1, data, mask_c0=None, mask_c1=None, defer_readback=False):
        """Launch the score / decision kernels and the asynchronous read-back of the match count; returns a handle for
        forward_end().  GPU work queued in between overlaps the host's wait for M (ops.dual_softmax_match_begin)."""
        if mask_c0 is not None or 'mask0' in data:
            raise NotImplementedError("padding masks (MegaDepth) are outside the FAR eval path")
        if self.training:
            raise NotImplementedError("training-time sampling (:199-240) is out of scope (SURVEY.md 8a a6)")
        scale = data['hw0_i'][0] / data['hw0_c'][0]
        if 'scale0' in data:
            raise NotImplementedError("per-image scale0/scale1 (MegaDepth resize) is outside the FAR eval path")
        return ops.dual_softmax_match_begin(feat_c0, feat_c1, tuple(data['hw0_c']), tuple(data['hw1_c']), self.thr,
                                            self.border_rm, self.temperature, scale, scale,
                                            return_conf_matrix=self.materialize_conf_matrix, engine=self.engine,
                                            defer_readback=defer_readback)

    def forward_end(self, handle, data):
        m = ops.dual_softmax_match_end(handle)
        data.update({'conf_matrix': m.get('conf_matrix'),
                     'b_ids': m['b_ids'], 'i_ids': m['i_ids'], 'j_ids': m['j_ids'],
                     'gt_mask': torch.zeros_like(m['mconf'], dtype=torch.bool),   # mconf == 0 never survives (:258)
                     'm_bids': m['b_ids'], 'mkpts0_c': m['mkpts0_c'], 'mkpts1_c': m['mkpts1_c'],
                     'mconf': m['mconf']})

    def forward(self, feat_c0, feat_c1, data, mask_c0=None, mask_c1=None):
        self.forward_end(self.forward_begin(feat_c0, feat_c1, data, mask_c0, mask_c1), data)
