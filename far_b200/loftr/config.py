"""Model config dicts (the lower-cased yacs config the reference hands to LoFTR, lightning_loftr.py:40-41)."""
import copy


def default_cfg():
    """mp3d_loftr/src/config/default.py:5-60, lower-cased, plus the argparse-poked FAR keys at their
    defaults (mp3d_loftr/test.py:162-223)."""
    return {
        'backbone_type': 'ResNetFPN', 'resolution': (8, 2), 'fine_window_size': 5, 'fine_concat_coarse_feat': True,
        'resnetfpn': {'initial_dim': 128, 'block_dims': [128, 196, 256]},
        'coarse': {'d_model': 256, 'd_ffn': 256, 'nhead': 8, 'layer_names': ['self', 'cross'] * 4,
                   'attention': 'linear', 'temp_bug_fix': True},
        'match_coarse': {'thr': 0.2, 'border_rm': 2, 'match_type': 'dual_softmax', 'dsmax_temperature': 0.1,
                         'skh_iters': 3, 'skh_init_bin_score': 1.0, 'skh_prefilter': False,
                         'train_coarse_percent': 0.2, 'train_pad_num_gt_min': 200, 'sparse_spvs': True},
        'fine': {'d_model': 128, 'd_ffn': 128, 'nhead': 8, 'layer_names': ['self', 'cross'], 'attention': 'linear'},
        'regress': {'d_model': 256, 'd_ffn': 256, 'nhead': 8, 'layer_names': ['self', 'cross'], 'attention': 'linear',
                    'temp_bug_fix': False, 'use_pos_embedding': False, 'regress_use_num_corres': False,
                    'save_mlp_feats': False, 'use_simple_moe': False, 'use_2wt': False, 'use_5050_weight': False,
                    'use_1wt': False, 'scale_8pt': False, 'save_gating_weights': False},
        'predict_translation_scale': False, 'regress_rt': False, 'regress_loftr_layers': 1,
        'from_saved_preds': None, 'save_preds': None, 'solver': 'ransac', 'use_many_ransac_thr': False,
        'fine_pred_steps': 1, 'training': False,
    }


def far_eval_cfg(thr=0.2, coarse_layers=3, regress_layers=1):
    """The FAR-LoFTR eval recipe of record: mp3d_loftr/scripts/eval_matterport.sh:18-37."""
    c = default_cfg()
    c['coarse']['layer_names'] = ['self', 'cross'] * coarse_layers
    c['match_coarse']['thr'] = thr
    c['regress'].update({'layer_names': ['self', 'cross'] * regress_layers, 'use_pos_embedding': True,
                         'regress_use_num_corres': True, 'use_simple_moe': True, 'use_2wt': True, 'scale_8pt': True,
                         'save_gating_weights': True})
    c.update({'regress_rt': True, 'regress_loftr_layers': regress_layers, 'solver': 'prior_ransac',
              'use_many_ransac_thr': True, 'fine_pred_steps': 2})
    return c


def upstream_loftr_cfg():
    """Pristine LoFTR default (mapfree_6dreg/etc/feature_matching_baselines/LoFTR/src/loftr/utils/cvpr_ds_config.py):
    4x(self,cross), temp_bug_fix False, no FAR head."""
    c = default_cfg()
    c['coarse']['temp_bug_fix'] = False
    c['match_coarse'].update({'skh_prefilter': True, 'train_coarse_percent': 0.4})
    return c


def clone(cfg):
    return copy.deepcopy(cfg)


class Cfg(dict):
    """Attribute-access dict with the layout of the reference's yacs node (upper-case keys): what `PL_LoFTR(config)`
    and `spvs_RT(data, config)` read (lightning_loftr.py:31-84, supervision.py:191-203)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _upper(d):
    return Cfg({k.upper(): (_upper(v) if isinstance(v, dict) else v) for k, v in d.items()})


def full_cfg(thr=0.2, coarse_layers=3, regress_layers=1):
    """The full eval config of the FAR-LoFTR recipe of record (mp3d_loftr/scripts/eval_matterport.sh over
    src/config/default.py, the keys the eval path reads): config.LOFTR.* (= far_eval_cfg, upper-cased), TRAINER.RANSAC_*,
    the harness switches of test.py:162-223."""
    c = Cfg()
    c.LOFTR = _upper(far_eval_cfg(thr, coarse_layers, regress_layers))
    c.TRAINER = Cfg({'RANSAC_PIXEL_THR': 0.5, 'RANSAC_CONF': 0.99999, 'N_VAL_PAIRS_TO_PLOT': 32, 'WORLD_SIZE': 1})
    c.update({'USE_CORRESPONDENCE_TRANSFORMER': False, 'USE_PRED_CORR': False, 'STRICT_FALSE': False, 'SAVE_PREDS': None,
              'NO_SAVE_PREDS': False, 'NO_SAVE_NUMCORR': False, 'SAVE_HARD_CORRES': False, 'SAVE_CORR': False,
              'EVAL_FIT_ONLY': False, 'PL_VERSION': '1.6.0', 'EVAL_SPLIT': 'test'})
    return c
