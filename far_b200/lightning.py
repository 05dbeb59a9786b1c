"""Eval-path mirror of the reference's Lightning wrapper, so the drop-in modules can be exercised through the same call
sequence `PL_LoFTR.test_step` runs (mp3d_loftr/src/lightning/lightning_loftr.py:31-84, 174-210, 325-421) without
pytorch_lightning (not in this image) and without the reference tree (absent on the GPU box):

    PL_LoFTR(config, pretrained_ckpt=None, profiler=None, dump_dir=None, split=None)   same constructor, `.matcher`
    test_step(batch, batch_idx, skip_eval=False)                                        same control flow and batch keys
    spvs_RT / compute_supervision_RT                                                    loftr/utils/supervision.py:184-240
    compute_pose_errors, relative_pose_error                                            utils/metrics.py:17-38, 198-303

The per-pair python loops and their host round trips are kept AS THE REFERENCE HAS THEM (this file is the reference
harness's shape, the batched device path is far_b200.pipeline.FarPosePipeline); the solver behind
`estimate_pose(...)` is far_b200.solver's GPU RANSAC shim.  `config` is the reference's yacs node or anything with the same
attribute layout (`far_b200.loftr.config.full_cfg()` builds one).  Training (`training_step`, losses, optimizers),
plotting and the correspondence-transformer ablation are outside the hot path and raise NotImplementedError."""
import os

import numpy as np
import torch
import torch.nn as nn

from .loftr import LoFTR
from .loftr.pose import pose_mean_6d, pose_std_6d, rotation_6d_to_matrix
from .solver import estimate_pose


def lower_config(cfg):
    """lightning_loftr.py / src/utils/misc.py `lower_config`: nested node -> plain dict with lower-case keys."""
    if not isinstance(cfg, dict):
        return cfg
    return {k.lower(): lower_config(v) for k, v in cfg.items()}


def relative_pose_error(T_0to1, R, t, ignore_gt_t_thr=0.0):
    """utils/metrics.py:17-38 (numpy)."""
    t_gt = T_0to1[:3, 3]
    n = np.linalg.norm(t) * np.linalg.norm(t_gt)
    t_err = np.rad2deg(np.arccos(np.clip(np.dot(t, t_gt) / n, -1.0, 1.0)))
    t_err = np.minimum(t_err, 180 - t_err)
    if np.linalg.norm(t_gt) < ignore_gt_t_thr:
        t_err = 0
    t_err_abs = np.linalg.norm(t - t_gt)
    cos = np.clip((np.trace(np.dot(R.T, T_0to1[:3, :3])) - 1) / 2, -1., 1.)
    return t_err, np.rad2deg(np.abs(np.arccos(cos))), t_err_abs


def spvs_RT(data, config):
    """loftr/utils/supervision.py:184-233: per-pair solver loop; like the reference it keeps the LAST pair's pose."""
    pixel_thr, conf = config.TRAINER.RANSAC_PIXEL_THR, config.TRAINER.RANSAC_CONF
    m_bids, pts0, pts1, K0, K1 = data['m_bids'], data['mkpts0_f'], data['mkpts1_f'], data['K0'], data['K1']
    priorRT = data['priorRT'] if (config.LOFTR.SOLVER == 'prior_ransac' and 'priorRT' in data) else None
    dev = pts0.device
    pred_rt = pred_e = None
    n_after = tight = ultra = 0
    mask = None
    for bs in range(K0.shape[0]):
        mask = m_bids == bs
        ret, n_after, tight, ultra = estimate_pose(pts0[mask], pts1[mask], K0[bs], K1[bs], pixel_thr, conf=conf,
                                                   translation_scale=data['translation_scale'],
                                                   solver=config.LOFTR.SOLVER, priorRT=priorRT)
        if ret is not None:
            pred_rt = torch.cat([ret[0], ret[1].unsqueeze(1)], axis=1)
            pred_e = ret[3]
        else:
            pred_rt = torch.cat([torch.eye(3), torch.zeros([3, 1])], axis=1).to(dev)
            pred_e = torch.eye(3).to(dev)
    data.update({"loftr_rt": pred_rt, "expec_rt": pred_rt, "expec_e": pred_e,
                 'num_correspondences_before_ransac': torch.tensor([int(mask.sum())]).to(dev),
                 'num_correspondences_after_ransac': n_after,
                 'num_correspondences': torch.tensor([int(n_after)]).to(dev),
                 'inliers_best_tight': torch.tensor([int(tight)]).to(dev),
                 'inliers_best_ultra_tight': torch.tensor([int(ultra)]).to(dev)})


def compute_supervision_RT(data, config):
    if data['dataset_name'][0].lower() in ['mp3d', 'interiornet_streetlearn']:
        spvs_RT(data, config)
    else:
        raise NotImplementedError


def compute_pose_errors(data, config):
    """utils/metrics.py:198-303 (the branches the FAR eval recipe reaches: regressed pose, or solver on the matches)."""
    pixel_thr, conf = config.TRAINER.RANSAC_PIXEL_THR, config.TRAINER.RANSAC_CONF
    data.update({'R_errs': [], 't_errs': [], 't_errs_abs': [], 'inliers': [], 'successful_fits': [], 'pred_R': [],
                 'pred_t': [], 'num_correspondences_before_ransac': [], 'num_correspondences_after_ransac': []})
    K0, K1 = data['K0'], data['K1']
    T_0to1 = data['T_0to1'].cpu().numpy()
    priorRT = None
    if config.LOFTR.SOLVER == 'prior_ransac' and 'priorRT' in data:
        priorRT = data['priorRT']
        if torch.is_tensor(priorRT):
            priorRT = priorRT.cpu().numpy()[0]
    for bs in range(K0.shape[0]):
        if 'regressed_rt' in data:
            rr = data['regressed_rt'].detach().cpu()
            R = rotation_6d_to_matrix(rr[:, 3:] * pose_std_6d[3:] + pose_mean_6d[3:])[0].numpy()
            t = rr[0, :3].numpy() * pose_std_6d[:3].numpy() + pose_mean_6d[:3].numpy()
            inliers = 0
            data['successful_fits'].append(0)
        elif 'mkpts0_f' in data:
            mask = data['m_bids'] == bs
            pts0, pts1 = data['mkpts0_f'], data['mkpts1_f']
            ret, n_after, _, _ = estimate_pose(pts0[mask], pts1[mask], K0[bs], K1[bs], pixel_thr, conf=conf,
                                               translation_scale=data['translation_scale'], solver=config.LOFTR.SOLVER,
                                               priorRT=priorRT)
            if ret is None:
                ret = (np.eye(3), np.random.rand(3) - .5, np.zeros(mask.shape[0]), np.eye(3))
                data['successful_fits'].append(0)
            else:
                ret = (ret[0].cpu().numpy(), ret[1].cpu().numpy(), ret[2], ret[3].cpu().numpy())
                data['successful_fits'].append(1)
                data['num_correspondences_before_ransac'].append(int(mask.sum()))
                data['num_correspondences_after_ransac'].append(n_after)
            R, t, inliers, _ = ret
        else:
            R, t, inliers = np.eye(3), np.random.rand(3) - .5, 0
            data['successful_fits'].append(0)
        t_err, R_err, t_err_abs = relative_pose_error(T_0to1[bs], R, t, ignore_gt_t_thr=0.0)
        data['pred_R'], data['pred_t'] = R, t
        data['R_errs'].append(R_err)
        data['t_errs'].append(t_err)
        data['t_errs_abs'].append(t_err_abs)
        data['inliers'].append(inliers)


class _PassThroughProfiler:
    class _Ctx:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def profile(self, name):
        return self._Ctx()


class PL_LoFTR(nn.Module):
    def __init__(self, config, pretrained_ckpt=None, profiler=None, dump_dir=None, split=None):
        super().__init__()
        self.config = config
        _config = lower_config(self.config)
        self.loftr_cfg = lower_config(_config['loftr'])
        self.profiler = profiler or _PassThroughProfiler()
        if getattr(config, 'USE_CORRESPONDENCE_TRANSFORMER', False):
            raise NotImplementedError("the correspondence-transformer ablation is outside the FAR path (SURVEY.md 2 row 13)")
        self.matcher = LoFTR(config=_config['loftr'])
        self.pretrained_ckpt = None
        if pretrained_ckpt:
            self.pretrained_ckpt = pretrained_ckpt
            state_dict = torch.load(pretrained_ckpt, map_location='cpu')['state_dict']
            self.matcher.load_state_dict(state_dict, strict=not getattr(config, 'STRICT_FALSE', False))
        self.dump_dir, self.split = dump_dir, split

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training is outside the per-pair pose hot path")

    def _compute_metrics(self, batch):
        """lightning_loftr.py:174-210 without the epipolar-error column (training-time GT geometry)."""
        compute_pose_errors(batch, self.config)
        rel_pair_names = list(zip(*batch['pair_names']))
        bs = batch['image0'].size(0)
        metrics = {'identifiers': ['#'.join(rel_pair_names[b]) for b in range(bs)],
                   'R_errs': batch['R_errs'], 't_errs': batch['t_errs'], 't_errs_abs': batch['t_errs_abs'],
                   'inliers': batch['inliers'], 'successful_fits': batch['successful_fits'],
                   'gt_R': batch['T_0to1'][:, :3, :3].cpu(),
                   'pred_R': torch.from_numpy(np.asarray(batch['pred_R'])).unsqueeze(0).cpu(),
                   'pred_t': torch.from_numpy(np.asarray(batch['pred_t'])).unsqueeze(0).cpu()}
        if 'gating_reg_weights' in batch:
            metrics['gating_reg_weights'] = batch['gating_reg_weights'].detach().cpu()
        return {'metrics': metrics}, rel_pair_names

    @torch.no_grad()
    def test_step(self, batch, batch_idx, skip_eval=False):
        cfg = self.config
        if cfg.LOFTR.FROM_SAVED_PREDS is None:
            with self.profiler.profile("LoFTR"):
                self.matcher(batch)
        if cfg.LOFTR.REGRESS_RT:
            if cfg.LOFTR.SOLVER == "prior_ransac_noprior" or \
                    (cfg.LOFTR.FROM_SAVED_PREDS is None and cfg.LOFTR.REGRESS.USE_SIMPLE_MOE):
                batch['translation_scale'] = None
                compute_supervision_RT(batch, cfg)
            for i in range(cfg.LOFTR.FINE_PRED_STEPS):
                with self.profiler.profile("LoFTR"):
                    self.matcher.forward_rt_prediction(batch)
                if i < cfg.LOFTR.FINE_PRED_STEPS - 1 and 'prior_ransac' in cfg.LOFTR.SOLVER:
                    compute_supervision_RT(batch, cfg)
        if skip_eval:
            return batch
        ret_dict, _ = self._compute_metrics(batch)
        parent = getattr(cfg, 'SAVE_PREDS', None)
        if parent is not None and not getattr(cfg, 'NO_SAVE_PREDS', False):   # :348-360, the 8pt-ViT prediction cache
            from .pred_cache import save_prediction
            nc = batch['num_correspondences_after_ransac'][0] if len(batch['num_correspondences_after_ransac']) > 0 \
                and not getattr(cfg, 'NO_SAVE_NUMCORR', False) else None
            save_prediction(parent, self.split, str(int(batch['pair_id'])), batch['pred_R'], batch['pred_t'], nc)
        return ret_dict
