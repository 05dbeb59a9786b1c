"""FinePreprocess (mp3d_loftr/src/loftr/loftr_module/fine_preprocess.py:7-59)."""
import torch
import torch.nn as nn

from .. import ops


class FinePreprocess(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.cat_c_feat = config['fine_concat_coarse_feat']
        self.W = self.config['fine_window_size']
        d_model_c = self.config['coarse']['d_model']
        d_model_f = self.config['fine']['d_model']
        self.d_model_f = d_model_f
        if not self.cat_c_feat:
            raise NotImplementedError("FINE_CONCAT_COARSE_FEAT=False is not used by any shipped config")
        self.down_proj = nn.Linear(d_model_c, d_model_f, bias=True)
        self.merge_feat = nn.Linear(2 * d_model_f, d_model_f, bias=True)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.kaiming_normal_(p, mode="fan_out", nonlinearity="relu")

    def forward(self, feat_f0, feat_f1, feat_c0, feat_c1, data):
        W = self.W
        stride = data['hw0_f'][0] // data['hw0_c'][0]
        data.update({'W': W})
        return ops.fine_preprocess(feat_f0, feat_f1, feat_c0, feat_c1, data['b_ids'], data['i_ids'], data['j_ids'],
                                   W, stride, data['hw0_c'][1], data['hw1_c'][1], self.down_proj.weight,
                                   self.down_proj.bias, self.merge_feat.weight, self.merge_feat.bias)
