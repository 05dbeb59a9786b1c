"""Import the UNMODIFIED reference (crockwell/far, mounted read-only at /root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  This file is used by `tests/golden/make_golden.py` (in the build
container, where /root/reference exists) to pin the restatement in `oracle/far_oracle.py` against the
reference's own code and to generate the committed fixtures under `tests/golden/`.  Nothing in the
product package (`far_b200/`) may import it, and nothing that runs on the GPU box may need it.

The reference has no executable parity pins of its own and needs third-party
packages that are not in this image.  The shims below are the ones SURVEY.md §8(c) lists:

  1. yacs.config.CfgNode            -> small attr-dict
  2. kornia dsnt.spatial_expectation2d / create_meshgrid (kornia==0.7.1 semantics restated)
  3. torch.Tensor.cuda / Module.cuda -> identity (the reference hard-codes .cuda() in ~40 places)
  4. pytorch_lightning.LightningModule -> nn.Module
  5. torchvision resnet18(pretrained=True) -> weights=None
  7. empty kornia.core / kornia.geometry modules so prior_ransac/cv_geometry.py imports
"""
import contextlib
import importlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("FAR_REFERENCE_ROOT", "/root/reference")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mp3d_loftr", "src", "loftr"))


# ----------------------------------------------------------------------------- shims
class _CfgNode(dict):
    """10-line stand-in for yacs.config.CfgNode (attribute access + clone)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        out = _CfgNode()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, _CfgNode) else (list(v) if isinstance(v, list) else v)
        return out

    def merge_from_file(self, path):  # not needed by the oracle
        raise NotImplementedError


def _kornia_create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=None):
    # kornia 0.7.1 kornia/utils/grid.py: xs = linspace(0, w-1, w); if normalized: (xs/(w-1)-0.5)*2
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    base = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)  # W x H x 2
    return base.permute(1, 0, 2).unsqueeze(0)  # 1 x H x W x 2


def _kornia_spatial_expectation2d(inp, normalized_coordinates=True):
    # kornia 0.7.1 kornia/geometry/subpix/dsnt.py
    b, c, h, w = inp.shape
    grid = _kornia_create_meshgrid(h, w, normalized_coordinates, inp.device).to(inp.dtype)
    pos_x = grid[..., 0].reshape(-1)
    pos_y = grid[..., 1].reshape(-1)
    flat = inp.view(b, c, -1)
    ex = torch.sum(pos_x * flat, -1, keepdim=True)
    ey = torch.sum(pos_y * flat, -1, keepdim=True)
    return torch.cat([ex, ey], -1).view(b, c, 2)


def _sampson_epipolar_distance(pts1, pts2, Fm, squared=True, eps=1e-8):
    # kornia 0.7.1 kornia/geometry/epipolar/_metrics.py
    if pts1.shape[-1] == 2:
        pts1 = torch.nn.functional.pad(pts1, [0, 1], value=1.0)
    if pts2.shape[-1] == 2:
        pts2 = torch.nn.functional.pad(pts2, [0, 1], value=1.0)
    F_t = Fm.transpose(-2, -1)
    line1_in_2 = pts1 @ F_t
    line2_in_1 = pts2 @ Fm
    numerator = (pts2 * line1_in_2).sum(dim=-1).pow(2)
    denominator = line1_in_2[..., :2].norm(2, dim=-1).pow(2) + line2_in_1[..., :2].norm(2, dim=-1).pow(2)
    out = numerator / denominator
    return out if squared else (out + eps).sqrt()


def _symmetrical_epipolar_distance(pts1, pts2, Fm, squared=True, eps=1e-8):
    if pts1.shape[-1] == 2:
        pts1 = torch.nn.functional.pad(pts1, [0, 1], value=1.0)
    if pts2.shape[-1] == 2:
        pts2 = torch.nn.functional.pad(pts2, [0, 1], value=1.0)
    F_t = Fm.transpose(-2, -1)
    line1_in_2 = pts1 @ F_t
    line2_in_1 = pts2 @ Fm
    numerator = (pts2 * line1_in_2).sum(dim=-1).pow(2)
    denom_inv = 1.0 / (line1_in_2[..., :2].norm(2, dim=-1).pow(2)) + 1.0 / (
        line2_in_1[..., :2].norm(2, dim=-1).pow(2))
    out = numerator * denom_inv
    return out if squared else (out + eps).sqrt()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


def install_shims():
    if "yacs" not in sys.modules:
        _mod("yacs")
        _mod("yacs.config", CfgNode=_CfgNode)
    if "kornia" not in sys.modules:
        anyfn = lambda *a, **k: None  # noqa: E731
        k = _mod("kornia")
        _mod("kornia.geometry", find_fundamental=anyfn, find_homography_dlt=anyfn,
             find_homography_dlt_iterated=anyfn, find_homography_lines_dlt=anyfn,
             find_homography_lines_dlt_iterated=anyfn,
             symmetrical_epipolar_distance=_symmetrical_epipolar_distance,
             epipolar=types.SimpleNamespace())
        _mod("kornia.geometry.subpix")
        dsnt = _mod("kornia.geometry.subpix.dsnt", spatial_expectation2d=_kornia_spatial_expectation2d)
        sys.modules["kornia.geometry.subpix"].dsnt = dsnt
        _mod("kornia.geometry.epipolar", sampson_epipolar_distance=_sampson_epipolar_distance,
             symmetrical_epipolar_distance=_symmetrical_epipolar_distance)
        _mod("kornia.geometry.epipolar.fundamental", fundamental_from_essential=anyfn)
        _mod("kornia.geometry.homography", line_segment_transfer_error_one_way=anyfn,
             oneway_transfer_error=anyfn, sample_is_valid_for_homography=anyfn)
        _mod("kornia.geometry.solvers", solve_cubic=anyfn)
        _mod("kornia.geometry.linalg", transform_points=anyfn)
        _mod("kornia.geometry.conversions", convert_points_to_homogeneous=anyfn, rotation_matrix_to_quaternion=anyfn,
             quaternion_to_rotation_matrix=anyfn, QuaternionCoeffOrder=types.SimpleNamespace(WXYZ=None))
        _mod("kornia.utils", create_meshgrid=_kornia_create_meshgrid)
        _mod("kornia.utils.grid", create_meshgrid=_kornia_create_meshgrid)
        _mod("kornia.utils.helpers", _torch_svd_cast=anyfn, safe_inverse_with_mask=anyfn,
             safe_solve_with_mask=anyfn, _extract_device_dtype=anyfn)
        _mod("kornia.core", Device=object, Module=torch.nn.Module, Tensor=torch.Tensor, zeros=torch.zeros,
             concatenate=torch.cat, ones_like=torch.ones_like, stack=torch.stack, where=torch.where,
             zeros_like=torch.zeros_like)
        _mod("kornia.core.check", KORNIA_CHECK_SHAPE=anyfn, KORNIA_CHECK_SAME_SHAPE=anyfn,
             KORNIA_CHECK_IS_TENSOR=anyfn, KORNIA_CHECK=anyfn)
        k.geometry = sys.modules["kornia.geometry"]
        k.geometry.solvers = sys.modules["kornia.geometry.solvers"]
    if "pytorch_lightning" not in sys.modules:
        _mod("pytorch_lightning", LightningModule=torch.nn.Module, seed_everything=lambda s: torch.manual_seed(s))
    # 3. .cuda() -> identity (CPU oracle)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


@contextlib.contextmanager
def _subproject(name):
    """Put one reference sub-project on sys.path; its top-level packages (`src`, `lib`, `third_party`)
    collide between sub-projects, so purge them on entry."""
    install_shims()
    root = os.path.join(REF_ROOT, name)
    for k in [k for k in sys.modules if k.split(".")[0] in ("src", "lib", "third_party", "config",
                                                          "essential", "cv_geometry", "torch_utils", "linalg",
                                                          "torch_version", "utils", "ransac")]:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        yield root
    finally:
        sys.path.remove(root)


def mp3d_eval_config(thr=0.0, coarse_layers=3, regress_layers=1):
    """The FAR-LoFTR config of record: mp3d_loftr/scripts/eval_matterport.sh:18-37 applied to
    src/config/default.py the way test.py:162-223 does, then lower-cased (lightning_loftr.py:40-41)."""
    cfg = {
        'backbone_type': 'ResNetFPN', 'resolution': (8, 2), 'fine_window_size': 5,
        'fine_concat_coarse_feat': True,
        'resnetfpn': {'initial_dim': 128, 'block_dims': [128, 196, 256]},
        'coarse': {'d_model': 256, 'd_ffn': 256, 'nhead': 8, 'layer_names': ['self', 'cross'] * coarse_layers,
                   'attention': 'linear', 'temp_bug_fix': True},
        'match_coarse': {'thr': thr, 'border_rm': 2, 'match_type': 'dual_softmax', 'dsmax_temperature': 0.1,
                         'skh_iters': 3, 'skh_init_bin_score': 1.0, 'skh_prefilter': False,
                         'train_coarse_percent': 0.2, 'train_pad_num_gt_min': 200, 'sparse_spvs': True},
        'fine': {'d_model': 128, 'd_ffn': 128, 'nhead': 8, 'layer_names': ['self', 'cross'], 'attention': 'linear'},
        'regress': {'d_model': 256, 'd_ffn': 256, 'nhead': 8, 'layer_names': ['self', 'cross'] * regress_layers,
                    'attention': 'linear', 'temp_bug_fix': False, 'use_pos_embedding': True,
                    'regress_use_num_corres': True, 'save_mlp_feats': False, 'use_simple_moe': True,
                    'use_2wt': True, 'use_5050_weight': False, 'use_1wt': False, 'scale_8pt': True,
                    'save_gating_weights': True},
        'predict_translation_scale': False, 'regress_rt': True, 'regress_loftr_layers': regress_layers,
        'from_saved_preds': None, 'save_preds': None, 'solver': 'prior_ransac', 'use_many_ransac_thr': True,
        'fine_pred_steps': 2, 'training': False,
    }
    return cfg


def load_mp3d():
    """Returns a namespace with the reference's mp3d_loftr classes/functions."""
    with _subproject("mp3d_loftr"):
        ns = types.SimpleNamespace()
        loftr_pkg = importlib.import_module("src.loftr")
        ns.LoFTR = loftr_pkg.LoFTR
        tr = importlib.import_module("src.loftr.loftr_module.transformer")
        ns.transformer = tr
        ns.LoFTREncoderLayer = tr.LoFTREncoderLayer
        ns.LocalFeatureTransformer = tr.LocalFeatureTransformer
        ns.CrossAttention = tr.CrossAttention
        ns.CrossBlock = tr.CrossBlock
        ns.LocalFeatureTransformerRegressor = tr.LocalFeatureTransformerRegressor
        ns.get_positional_encodings = tr.get_positional_encodings
        ns.LinearAttention = importlib.import_module("src.loftr.loftr_module.linear_attention").LinearAttention
        ns.FinePreprocess = importlib.import_module("src.loftr.loftr_module.fine_preprocess").FinePreprocess
        ns.CoarseMatching = importlib.import_module("src.loftr.utils.coarse_matching").CoarseMatching
        ns.FineMatching = importlib.import_module("src.loftr.utils.fine_matching").FineMatching
        ns.PositionEncodingSine = importlib.import_module("src.loftr.utils.position_encoding").PositionEncodingSine
        ns.loss = importlib.import_module("src.losses.loftr_loss")
        ns.backbone = importlib.import_module("src.loftr.backbone")
        return ns


def load_prior_ransac():
    """Reference solver functions: run_8point, decompose_essential_matrix, motion_from_essential."""
    with _subproject("mp3d_loftr") as root:
        pr = os.path.join(root, "third_party", "prior_ransac")
        sys.path.insert(0, pr)
        try:
            ns = types.SimpleNamespace()
            cvg = importlib.import_module("cv_geometry")
            ess = importlib.import_module("essential")
            ns.run_8point = cvg.run_8point
            ns.normalize_points = cvg.normalize_points
            ns.decompose_essential_matrix = ess.decompose_essential_matrix
            ns.motion_from_essential = ess.motion_from_essential
            return ns
        finally:
            sys.path.remove(pr)


def vit8pt_args():
    """argparse.Namespace of the 8pt-ViT eval recipe (interiornetStreetlearn_8ptVit/scripts + train.py:403-445)."""
    return types.SimpleNamespace(pool_size=60, fc_hidden_size=512, use_loftr_gating=True, use_normalized_6d=True,
                                 fusion_transformer=True, transformer_depth=6,
                                 T_pose=torch.tensor([[0., 0., 1.]]))


def load_vit8pt():
    """Reference ViTEss class (interiornetStreetlearn_8ptVit/src/model.py) with resnet18(pretrained=True) -> weights=None."""
    with _subproject("interiornetStreetlearn_8ptVit"):
        import torchvision.models as tvm
        orig = tvm.resnet18
        tvm.resnet18 = lambda pretrained=False, **kw: orig(weights=None)
        try:
            ns = types.SimpleNamespace()
            mod = importlib.import_module("src.model")
            ns.ViTEss = mod.ViTEss
            ns.vt = importlib.import_module("src.modules.vision_transformer")
            ns.resnet18_patch = (tvm, orig)
            return ns
        finally:
            pass


def _attr_cfg(d):
    out = _CfgNode()
    for k, v in d.items():
        out[k] = _attr_cfg(v) if isinstance(v, dict) else v
    return out


def mapfree_config():
    """config/regression/mapfree/rot6d_trans_with_loftr.yaml merged over config/default.py (the keys the model reads)."""
    return _attr_cfg({
        'MODEL': 'Regression',
        'ENCODER': {'TYPE': 'ResUNet', 'BLOCK_TYPE': 1, 'NUM_BLOCKS': '3-3-3', 'NOT_CONCAT': False, 'NUM_OUT_LAYERS': 32},
        'AGGREGATOR': {'TYPE': 'CorrelationVolumeWarping', 'POSITION_ENCODER': True, 'POSITION_ENCODER_IM1': None,
                       'MAX_SCORE_CHANNEL': True, 'NORMALISE_DOT': False, 'RESIDUAL_ATT': False, 'CV_OUTLAYERS': 0,
                       'CV_HALF_CHANNELS': False, 'UPSAMPLE_POS_ENC': 0, 'DUSTBIN': False},
        'HEAD': {'TYPE': 'DirectDeepResBlockMLP', 'ADD_BASIS': True, 'NUM_PTS': 6, 'AVG_POOL': True, 'BATCH_NORM': True,
                 'SEPARATE_SCALE': True},
        'TRAINING': {'ROT_LOSS': 'rot_6d_loss', 'TRANS_LOSS': 'trans_unnormalized_loss', 'LAMBDA': 1.},
        'BACKPROJECT_ANCHORS': False,
        'DATASET': {'HEIGHT': 360, 'WIDTH': 270},
        'SOLVER': {'EMAT_RANSAC': {'PIX_THRESHOLD': 2.0, 'SCALE_THRESHOLD': 0.1, 'CONFIDENCE': 0.9999}},
    })


def load_mapfree():
    """Reference map-free classes: RegressionModel (mapfree_6dreg/lib/models/regression/model.py), the pristine
    upstream LoFTR it embeds (etc/feature_matching_baselines/LoFTR/src/loftr) and its default_cfg.  torch.load of the
    (absent) outdoor_ot.ckpt is answered with an empty state dict (strict=False at model.py:105)."""
    with _subproject("mapfree_6dreg") as root:
        for k in [k for k in sys.modules if k.split(".")[0] in ("etc",)]:
            del sys.modules[k]
        pr = os.path.join(root, "third_party", "prior_ransac")
        sys.path.insert(0, pr)
        orig_load = torch.load
        torch.load = lambda *a, **k: {'state_dict': {}}
        try:
            ns = types.SimpleNamespace()
            mod = importlib.import_module("lib.models.regression.model")
            ns.model_module = mod
            ns.RegressionModel = mod.RegressionModel
            lo = importlib.import_module("etc.feature_matching_baselines.LoFTR.src.loftr")
            ns.UpstreamLoFTR, ns.upstream_default_cfg = lo.LoFTR, lo.default_cfg
            ns.build = lambda **kw: _build_mapfree(mod, kw)
            return ns
        finally:
            sys.path.remove(pr)
            ns.restore = lambda: setattr(torch, "load", orig_load)


def _build_mapfree(mod, kw):
    m = mod.RegressionModel(mapfree_config(), **kw)
    return m.eval()
