"""FAR-LoFTR top module with the reference's interface (mp3d_loftr/src/loftr/loftr.py:14-211):
`LoFTR(config)`, `forward(data, train=False)`, and the sub-steps PL_LoFTR calls directly
(`forward_feature_extraction`, `forward_correspondence_prediction`, `forward_rt_prediction`).  Results are
returned by mutating `data`, with the same keys.

Differences that are deliberate and documented:
  * a batch of N pairs is N independent B=1 evaluations (the reference head is batch-1 only);
    with N == 1 every output has the reference's shape (`regressed_rt [1,9]`, `priorRT` numpy [3,4]).
  * `data['conf_matrix']` is None unless config['match_coarse']['materialize_conf_matrix'] is set.
"""
import torch
import torch.nn as nn

from .backbone import build_backbone
from .position_encoding import PositionEncodingSine
from .transformer import LocalFeatureTransformer, LocalFeatureTransformerRegressor
from .fine_preprocess import FinePreprocess
from .coarse_matching import CoarseMatching
from .fine_matching import FineMatching
from .pose import compute_normalized_6d, rotation_6d_to_matrix, pose_mean_6d, pose_std_6d


class LoFTR(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if self.config.get('from_saved_preds') is None:
            self.backbone = build_backbone(config)
            self.pos_encoding = PositionEncodingSine(config['coarse']['d_model'],
                                                     temp_bug_fix=config['coarse']['temp_bug_fix'])
            self.loftr_coarse = LocalFeatureTransformer(config['coarse'])
            self.coarse_matching = CoarseMatching(config['match_coarse'])
            self.fine_preprocess = FinePreprocess(config)
            self.loftr_fine = LocalFeatureTransformer(config["fine"])
            self.fine_matching = FineMatching(config)
            if self.config.get('predict_translation_scale'):
                raise NotImplementedError("predict_translation_scale is off in every shipped FAR recipe")
        if self.config.get('regress_rt'):
            self.loftr_regress = LocalFeatureTransformerRegressor(config)

    # ------------------------------------------------------------------ 1. CNN (cuDNN, channels_last)
    def forward_feature_extraction(self, data):
        data.update({'bs': data['image0'].size(0),
                     'hw0_i': data['image0'].shape[2:], 'hw1_i': data['image1'].shape[2:]})
        if data['hw0_i'] == data['hw1_i']:
            feats_c, feats_f = self.backbone(torch.cat([data['image0'], data['image1']], dim=0))
            (feat_c0, feat_c1), (feat_f0, feat_f1) = feats_c.split(data['bs']), feats_f.split(data['bs'])
        else:
            feats_c = None
            (feat_c0, feat_f0), (feat_c1, feat_f1) = self.backbone(data['image0']), self.backbone(data['image1'])
        data.update({'hw0_c': feat_c0.shape[2:], 'hw1_c': feat_c1.shape[2:],
                     'hw0_f': feat_f0.shape[2:], 'hw1_f': feat_f1.shape[2:]})
        data.update({'featmap0': feat_c0, 'featmap1': feat_c1, 'featmap_f0': feat_f0, 'featmap_f1': feat_f1,
                     'feats_c': feats_c})

    # ------------------------------------------------------------------ 2-5. transformer + matching (CUDA kernels)
    def forward_coarse(self, data, defer_readback=False):
        """Positional encoding -> coarse LoFTR -> score / decision kernels of the coarse matcher.  Stops before the
        reference's host sync on the match count (coarse_matching.py:193): returns the handle forward_fine() completes,
        so a caller can queue work that only needs the coarse features (the FAR head trunk) in between."""
        if 'mask0' in data:
            raise NotImplementedError("padding masks (MegaDepth) are outside the FAR eval path")
        feat_c0 = self.pos_encoding.forward_flatten(data['featmap0'])   # [N, HW, C]
        feat_c1 = self.pos_encoding.forward_flatten(data['featmap1'])
        feat_c0, feat_c1 = self.loftr_coarse(feat_c0, feat_c1)
        handle = self.coarse_matching.forward_begin(feat_c0, feat_c1, data, defer_readback=defer_readback)
        data.update({'featmap0': feat_c0, 'featmap1': feat_c1, 'mask_c0': None, 'mask_c1': None,
                     'translation_scale': None})
        return handle

    def forward_fine(self, data, handle, train=False):
        feat_c0, feat_c1 = data['featmap0'], data['featmap1']
        self.coarse_matching.forward_end(handle, data)
        ff0, ff1 = self.fine_preprocess(data['featmap_f0'], data['featmap_f1'], feat_c0, feat_c1, data)
        if ff0.size(0) != 0:
            ff0, ff1 = self.loftr_fine(ff0, ff1)
        self.fine_matching(ff0, ff1, data, train=train)

    def forward_correspondence_prediction(self, data, train=False):
        self.forward_fine(data, self.forward_coarse(data), train=train)

    # ------------------------------------------------------------------ 6. FAR head
    def preprocess_helper(self, data):
        """Solver pose -> normalised 9-D (+ counters/500) for each pair (loftr.py:137-171).  Accepts the reference's
        B=1 shapes ([3,4] / [1,3,4], counters of shape [1]) and batched [N,3,4] / [N]."""
        feat_c0, feat_c1 = data['featmap0'], data['featmap1']
        loftr_preds_6d = inv_loftr_preds_6d = None
        rc = self.config['regress']
        if rc['use_simple_moe']:
            rt = data['loftr_rt'].detach()
            rt = rt.reshape(-1, 3, 4)
            dev = feat_c0.device
            rt = rt.to(dev)
            loftr_preds_6d = compute_normalized_6d(rt.float())
            bottom = torch.tensor([0, 0, 0, 1.], dtype=rt.dtype, device=dev).expand(rt.shape[0], 1, 4)
            inv = torch.linalg.inv(torch.cat([rt, bottom], dim=1))[:, :3, :4]
            inv_loftr_preds_6d = compute_normalized_6d(inv).float()

            def col(key):
                return data[key].detach().float().reshape(-1, 1).to(dev) / 500

            extra = []
            if rc['regress_use_num_corres']:
                extra.append(col('num_correspondences'))
            if self.config['use_many_ransac_thr']:
                extra += [col('num_correspondences_before_ransac'), col('inliers_best_tight'),
                          col('inliers_best_ultra_tight')]
            if extra:
                e = torch.cat(extra, dim=-1)
                loftr_preds_6d = torch.cat([loftr_preds_6d, e], dim=-1)
                inv_loftr_preds_6d = torch.cat([inv_loftr_preds_6d, e], dim=-1)
        return feat_c0, feat_c1, None, None, loftr_preds_6d, inv_loftr_preds_6d

    def head_trunk(self, data, feat_c0=None, feat_c1=None):
        """The solver-independent part of the FAR head on data['featmap0/1'] (post coarse transformer), cached in `data`
        for the current forward when config['regress']['reuse_trunk'] (default).  May be called right after
        forward_coarse(): it does not need the matches."""
        feat_c0 = data['featmap0'] if feat_c0 is None else feat_c0
        feat_c1 = data['featmap1'] if feat_c1 is None else feat_c1
        reuse = self.config['regress'].get('reuse_trunk', True)
        key = (feat_c0.data_ptr(), feat_c1.data_ptr(), feat_c0._version, feat_c1._version, tuple(feat_c0.shape))
        cached = data.get('_far_head_trunk') if reuse else None
        if cached is not None and cached[0] == key:
            return cached[1]
        trunk = self.loftr_regress.forward_trunk(feat_c0, feat_c1)
        if reuse:
            data['_far_head_trunk'] = (key, trunk, feat_c0, feat_c1)  # holds the maps: storage cannot be recycled
        return trunk

    def forward_rt_prediction(self, data):
        if not self.config['regress_rt']:
            return
        feat_c0, feat_c1, _, _, lp, ilp = self.preprocess_helper(data)
        # The head's trunk (regress LoFTR layers -> EMM -> encoder -> regressed pose) depends only on the two feature
        # maps; PL_LoFTR.test_step invokes the head fine_pred_steps = 2 times on the SAME maps with a different solver
        # prediction (lightning_loftr.py:338-346).  Within one forward (one `data` dict) the trunk is evaluated once and
        # reused; outputs are identical to re-evaluating it (tests/test_gpu_parity.py::test_head_trunk_reuse).
        # config['regress']['reuse_trunk'] = False restores the literal re-evaluation.
        trunk = self.head_trunk(data, feat_c0, feat_c1)
        pred_RT, mlp_features, pred_RT_wt = self.loftr_regress.forward_gate(trunk, loftr_preds=lp, inv_loftr_preds=ilp)
        data.update({'regressed_rt': pred_RT, 'expec_rt': pred_RT[0]})
        if self.config['regress']['save_mlp_feats']:
            data.update({'mlp_feats': mlp_features})
        if self.config['regress']['save_gating_weights']:
            data.update({'gating_reg_weights': pred_RT_wt})
        if self.config['solver'] == 'prior_ransac':  # loftr.py:187-192
            dev = pred_RT.device
            rr = pred_RT.detach()
            R = rotation_6d_to_matrix(rr[:, 3:] * pose_std_6d[3:].to(dev) + pose_mean_6d[3:].to(dev))
            t = rr[:, :3] * pose_std_6d[:3].to(dev) + pose_mean_6d[:3].to(dev)
            prior = torch.cat([R, t.unsqueeze(-1)], dim=-1)
            data['priorRT_device'] = prior          # [N,3,4] on the device: what the batched GPU RANSAC round reads
            # The reference hands spvs_RT a numpy [3,4] (a blocking D2H copy per head invocation).  Kept for the
            # reference harness (PL_LoFTR.test_step -> spvs_RT); FarPosePipeline sets config['regress']
            # ['prior_rt_on_device'] and never leaves the device.
            if not self.config['regress'].get('prior_rt_on_device', False):
                prior = prior.cpu().numpy()
                data.update({'priorRT': prior[0] if prior.shape[0] == 1 else prior})

    def forward(self, data, train=False):
        self.forward_feature_extraction(data)
        self.forward_correspondence_prediction(data, train=train)

    def load_state_dict(self, state_dict, *args, **kwargs):
        for k in list(state_dict.keys()):
            if k.startswith('matcher.'):
                state_dict[k.replace('matcher.', '', 1)] = state_dict.pop(k)
        return super().load_state_dict(state_dict, *args, **kwargs)
