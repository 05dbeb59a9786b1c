"""Map-free FAR model (mapfree_6dreg/lib/models/regression/model.py `RegressionModel`, BASELINE configs[3]):

    per pair           upstream LoFTR matcher on [1,720,544] (6120 coarse tokens, 8 coarse layers) -> essential-matrix
                       solver (model.py:241-276, pose_solver.py:30-97); with use_prior a second, prior-guided round
    ResUNet x2         [B,3,360,270] -> [B,32,92,68]                 (encoder/resunet.py:41-128; cuDNN, channels_last)
    aggregator         CorrelationVolumeWarping: softmax(6256^2) warp + soft position + max score -> [B,67,92,68]
                       (aggregator.py:42-116)  -- far_corr_volume_warp: flash-style tcgen05 kernel, volume never in HBM
    head               DirectDeepResBlockMLP(full_forward_pass=False) -> [B,256,12,9]   (head.py:27-55,248-281; cuDNN)
    transformer        nn.TransformerEncoder(6 x post-norm, d 256, 8 heads, 108 tokens) (model.py:57-61,288-291)
                       -- far_linear / far_softmax_attention / far_layernorm_pre
    regression_mlp     gated fusion with the solver pose (model.py:198-233)  -- far_linear (split-K)

Same class / parameter names as the reference (`encoder.*`, `transformer.layers.{i}.*`, `head.*`, `matcher.*`,
`pose_regressor.*`, `moe_predictor.*`), so a RegressionModel checkpoint loads; `forward(data) -> (R6d [B,6], t [B,3])`
with the reference's data-dict keys.  Deliberate differences (outputs identical): the B matcher evaluations run as one
batch instead of a python loop; with use_prior the matcher and the image branch (encoder .. transformer), which do not
depend on the prior, are evaluated once and reused by the second loop (the reference recomputes both, :240-291).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

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


# ---------------------------------------------------------------------------------------------------------------------
# encoder / head blocks (cuDNN, channels_last; eval-mode BatchNorm folded into the producing convolution where the BN
# is that convolution's only consumer)
def _fold(conv, bn):
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = (conv.weight * scale[:, None, None, None]).contiguous(memory_format=torch.channels_last)
    b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    return w, ((b0 - bn.running_mean) * scale + bn.bias).contiguous()


def _affine(bn):
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return scale.contiguous(), (bn.bias - bn.running_mean * scale).contiguous()


class _FoldCache:
    """Per-module cache of folded weights, invalidated when any parameter / buffer changes (version counters)."""

    def _cached(self, build):
        key = tuple(int(t._version) for t in list(self.parameters()) + list(self.buffers())) + \
            tuple(t.data_ptr() for t in self.parameters())
        c = getattr(self, "_fold_cache", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                c = (key, build())
            self._fold_cache = c
        return c[1]

    def _fast(self, x):
        # `force_eager` (set by bench.py's gpu_eager_baseline leg) keeps the module on the reference's plain op sequence
        return x.is_cuda and not self.training and not torch.is_grad_enabled() and not getattr(self, 'force_eager', False)


def _bn_relu(x, bn, aff):
    """relu(bn(x)) out of place (pre-activation: x itself is still the identity shortcut)."""
    if aff is not None and x.shape[1] % 4 == 0 and x.is_contiguous(memory_format=torch.channels_last):
        return ops.scale_shift_act(x, aff[0], aff[1], 0.0)
    return F.relu(bn(x))


class PreActBlock(nn.Module, _FoldCache):
    """encoder/preact.py:15-42."""
    expansion = 1

    def __init__(self, in_planes, planes, stride=1, bn=True):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(in_planes) if bn else nn.Identity()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes) if bn else nn.Identity()
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride,
                                                    bias=False))

    def forward(self, x):
        if self._fast(x) and isinstance(self.bn2, nn.BatchNorm2d):
            f = self._cached(lambda: {"a1": _affine(self.bn1), "c1": _fold(self.conv1, self.bn2)})
            out = _bn_relu(x, self.bn1, f["a1"])
            shortcut = self.shortcut(out) if hasattr(self, 'shortcut') else x
            out = torch.cudnn_convolution_relu(out, f["c1"][0], f["c1"][1], self.conv1.stride, (1, 1), (1, 1), 1)
            out = self.conv2(out)
            return out.add_(shortcut)
        out = F.relu(self.bn1(x))
        shortcut = self.shortcut(out) if hasattr(self, 'shortcut') else x
        out = self.conv1(out)
        out = self.conv2(F.relu(self.bn2(out)))
        out += shortcut
        return out


class PreActBottleneck(nn.Module, _FoldCache):
    """encoder/preact.py:45-75."""
    expansion = 4

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(in_planes)
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, self.expansion * planes, kernel_size=1, bias=False)
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride,
                                                    bias=False))

    def forward(self, x):
        if self._fast(x):
            f = self._cached(lambda: {"a1": _affine(self.bn1), "c1": _fold(self.conv1, self.bn2),
                                      "c2": _fold(self.conv2, self.bn3)})
            out = _bn_relu(x, self.bn1, f["a1"])
            shortcut = self.shortcut(out) if hasattr(self, 'shortcut') else x
            out = torch.cudnn_convolution_relu(out, f["c1"][0], f["c1"][1], (1, 1), (0, 0), (1, 1), 1)
            out = torch.cudnn_convolution_relu(out, f["c2"][0], f["c2"][1], self.conv2.stride, (1, 1), (1, 1), 1)
            out = self.conv3(out)
            return out.add_(shortcut)
        out = F.relu(self.bn1(x))
        shortcut = self.shortcut(out) if hasattr(self, 'shortcut') else x
        out = self.conv1(out)
        out = self.conv2(F.relu(self.bn2(out)))
        out = self.conv3(F.relu(self.bn3(out)))
        out += shortcut
        return out


class conv(nn.Module, _FoldCache):
    """encoder/resunet.py:17-28: Conv2d -> BatchNorm2d -> ELU."""

    def __init__(self, num_in_layers, num_out_layers, kernel_size, stride):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv = nn.Conv2d(num_in_layers, num_out_layers, kernel_size=kernel_size, stride=stride,
                              padding=(self.kernel_size - 1) // 2)
        self.normalize = nn.BatchNorm2d(num_out_layers)

    def forward(self, x):
        if self._fast(x):
            w, b = self._cached(lambda: _fold(self.conv, self.normalize))
            return F.elu(F.conv2d(x, w, b, self.conv.stride, self.conv.padding), inplace=True)
        return F.elu(self.normalize(self.conv(x)), inplace=True)


class upconv(nn.Module):
    """encoder/resunet.py:31-40: bilinear x2 (align_corners) -> conv."""

    def __init__(self, num_in_layers, num_out_layers, kernel_size, scale):
        super().__init__()
        self.scale = scale
        self.conv1 = conv(num_in_layers, num_out_layers, kernel_size, 1)

    def forward(self, x):
        if x.is_cuda and self.scale == 2 and x.shape[1] % 4 == 0 and not torch.is_grad_enabled() and \
                not getattr(self, 'force_eager', False):
            x = ops.upsample2x_add(x.contiguous(memory_format=torch.channels_last))
        else:
            x = F.interpolate(x, scale_factor=self.scale, mode='bilinear', align_corners=True)
        return self.conv1(x)


class ResUNet(nn.Module):
    """encoder/resunet.py:41-128.  cfg: object / dict with BLOCK_TYPE, NUM_BLOCKS, NOT_CONCAT, NUM_OUT_LAYERS."""

    def __init__(self, cfgmodel, num_in_layers=3):
        super().__init__()
        get = (lambda k, d=None: cfgmodel.get(k, d)) if isinstance(cfgmodel, dict) else \
            (lambda k, d=None: getattr(cfgmodel, k, d))
        filters = [256, 512, 1024, 2048]
        self.in_planes = 64
        self.firstconv = nn.Conv2d(num_in_layers, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.firstbn = nn.BatchNorm2d(64)
        self.firstrelu = nn.ReLU(inplace=True)
        self.firstmaxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        block = [PreActBlock, PreActBottleneck][get('BLOCK_TYPE')]
        num_blocks = [int(x) for x in get('NUM_BLOCKS').strip().split("-")]
        self.encoder1 = self._make_layer(block, 64, num_blocks[0], stride=1)
        self.encoder2 = self._make_layer(block, 128, num_blocks[1], stride=2)
        self.encoder3 = self._make_layer(block, 256, num_blocks[2], stride=2)
        self.not_concat = bool(get('NOT_CONCAT', False))
        self.upconv4 = upconv(filters[2], 512, 3, 2)
        self.iconv4 = conv((filters[1] + 512) if not self.not_concat else 512, 512, 3, 1)
        self.upconv3 = upconv(512, 256, 3, 2)
        self.iconv3 = conv((filters[0] + 256) if not self.not_concat else 256, 256, 3, 1)
        self.num_out_layers = get('NUM_OUT_LAYERS', 128)
        self.outconv = conv(256, self.num_out_layers, 1, 1)

    def _make_layer(self, block, planes, num_blocks, stride):
        layers = []
        for s in [stride] + [1] * (num_blocks - 1):
            layers.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*layers)

    @staticmethod
    def skipconnect(x1, x2):
        dY, dX = x2.size(2) - x1.size(2), x2.size(3) - x1.size(3)
        x1 = F.pad(x1, (dX // 2, dX - dX // 2, dY // 2, dY - dY // 2))
        return torch.cat([x2, x1], dim=1)

    def forward(self, x):
        if x.is_cuda:
            if not self.firstconv.weight.is_contiguous(memory_format=torch.channels_last):
                self.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)
        x1 = self.firstmaxpool(self.firstrelu(self.firstbn(self.firstconv(x))))
        x2 = self.encoder1(x1)
        x3 = self.encoder2(x2)
        x4 = self.encoder3(x3)
        x = self.upconv4(x4)
        if not self.not_concat:
            x = self.skipconnect(x3, x)
        x = self.iconv4(x)
        x = self.upconv3(x)
        if not self.not_concat:
            x = self.skipconnect(x2, x)
        x = self.iconv3(x)
        return self.outconv(x)


class CorrelationVolumeWarping(nn.Module):
    """aggregator.py:6-116 for the shipped recipe's flags (POSITION_ENCODER, MAX_SCORE_CHANNEL; no dustbin /
    normalisation / CV layers / IM1 encoder / half channels): one flash-style tcgen05 kernel."""

    def __init__(self, cfg, volume_channels):
        super().__init__()
        get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
        unsupported = [k for k in ("POSITION_ENCODER_IM1", "CV_OUTLAYERS", "CV_HALF_CHANNELS", "UPSAMPLE_POS_ENC", "DUSTBIN",
                                   "NORMALISE_DOT") if get(k)]
        if unsupported or not get("POSITION_ENCODER") or not get("MAX_SCORE_CHANNEL"):
            raise NotImplementedError("CorrelationVolumeWarping: only the FAR map-free recipe (POSITION_ENCODER + "
                                      f"MAX_SCORE_CHANNEL) is on the path; got {unsupported}")
        self.num_out_layers = 2 * volume_channels + 3

    def forward(self, vol0, vol1):
        assert vol0.shape == vol1.shape, 'Feature volumes shape must match'
        return ops.corr_volume_warp(vol0, vol1)


class DirectDeepResBlockMLP(nn.Module):
    """head.py:27-55 (DeepResBlock) + :248-281 with full_forward_pass=False, the only mode RegressionModel builds
    (model.py:71): three stride-2 PreActBlocks; returns (None, None, x3 [B,256,12,9])."""

    def __init__(self, cfg, in_channels, full_forward_pass=False):
        super().__init__()
        if full_forward_pass:
            raise NotImplementedError("RegressionModel builds the head with full_forward_pass=False (model.py:71)")
        head = cfg['HEAD'] if isinstance(cfg, dict) else cfg.HEAD
        bn = head.get('BATCH_NORM', True) if isinstance(head, dict) else getattr(head, 'BATCH_NORM', True)
        self.resblock1 = PreActBlock(in_channels, 64, stride=2, bn=bn)
        self.resblock2 = PreActBlock(64, 128, stride=2, bn=bn)
        self.resblock3 = PreActBlock(128, 256, stride=2, bn=bn)
        self.full_forward_pass = False

    def forward(self, feature_volume, data=None):
        if feature_volume.is_cuda:
            if not self.resblock1.conv1.weight.is_contiguous(memory_format=torch.channels_last):
                self.to(memory_format=torch.channels_last)
            feature_volume = feature_volume.contiguous(memory_format=torch.channels_last)
        x3 = self.resblock3(self.resblock2(self.resblock1(feature_volume)))
        return None, None, x3


def transformer_encoder(enc, x):
    """nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model, nhead), num_layers) in eval mode on token-major
    x [B, S, E] (torch defaults: post-norm, ReLU, in_proj / out_proj with bias; model.py:57-61, 288-291) through the
    sm_100a kernels: in_proj GEMM -> far_softmax_attention -> out_proj GEMM -> LN(x + .) -> linear1 (ReLU) -> linear2
    -> LN(x + .).  `enc` only holds the parameters (reference names transformer.layers.{i}.*)."""
    for layer in enc.layers:
        at = layer.self_attn
        h = at.num_heads
        d = at.embed_dim // h
        qkv = ops.linear(x, at.in_proj_weight, at.in_proj_bias)                        # [B,S,3E] == [B,S,3,h,d]
        a = ops.softmax_attention(qkv, h, 1.0 / math.sqrt(d))
        a = ops.linear(a, at.out_proj.weight, at.out_proj.bias)
        x = ops.layernorm(a, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, pre_add=x.reshape(-1, x.shape[-1]))
        ff = ops.linear(ops.linear(x, layer.linear1.weight, layer.linear1.bias, ACT_RELU), layer.linear2.weight,
                        layer.linear2.bias)
        x = ops.layernorm(ff, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, pre_add=x.reshape(-1, x.shape[-1]))
    return x


def rotation_6d_to_matrix(d6):
    """model.py:25-31."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def mapfree_cfg():
    """config/regression/mapfree/rot6d_trans_with_loftr.yaml over config/default.py, the keys the model reads."""
    return {'ENCODER': {'TYPE': 'ResUNet', 'BLOCK_TYPE': 1, 'NUM_BLOCKS': '3-3-3', 'NOT_CONCAT': False,
                        'NUM_OUT_LAYERS': 32},
            'AGGREGATOR': {'TYPE': 'CorrelationVolumeWarping', 'POSITION_ENCODER': True, 'POSITION_ENCODER_IM1': None,
                           'MAX_SCORE_CHANNEL': True, 'NORMALISE_DOT': False, 'RESIDUAL_ATT': False, 'CV_OUTLAYERS': 0,
                           'CV_HALF_CHANNELS': False, 'UPSAMPLE_POS_ENC': 0, 'DUSTBIN': False},
            'HEAD': {'TYPE': 'DirectDeepResBlockMLP', 'ADD_BASIS': True, 'AVG_POOL': True, 'BATCH_NORM': True},
            'DATASET': {'HEIGHT': 360, 'WIDTH': 270},
            'SOLVER': {'EMAT_RANSAC': {'PIX_THRESHOLD': 2.0, 'SCALE_THRESHOLD': 0.1, 'CONFIDENCE': 0.9999}}}


def _cfg_get(cfg, *path):
    for k in path:
        cfg = cfg[k] if isinstance(cfg, dict) else getattr(cfg, k)
    return cfg


class RegressionModel(RegressionHead):
    """Drop-in for mapfree_6dreg/lib/models/regression/model.py:33-308 (inference): same constructor keywords, same
    parameter names, `forward(data) -> (R6d [B,6], t [B,3])`, writes data['R'], ['t'], ['loftr_rt'], ['inliers'].
    data: image0/1 [B,1,720,544] (matcher), image0_reg/1_reg [B,3,360,270], K_color0/1 [B,3,3]."""

    def __init__(self, cfg=None, use_loftr_preds=False, use_superglue_preds=False, ckpt_path=None, not_strict=False,
                 inference=False, use_vanilla_transformer=False, d_model=32, max_steps=200_000, use_prior=False,
                 matcher_config=None):
        nn.Module.__init__(self)
        cfg = cfg if cfg is not None else mapfree_cfg()
        self.cfg = cfg
        if use_superglue_preds:
            raise NotImplementedError("SuperGlue is an alternative matcher outside the FAR path (SURVEY.md 2 row 25)")
        if _cfg_get(cfg, 'ENCODER', 'TYPE') != 'ResUNet' or _cfg_get(cfg, 'AGGREGATOR', 'TYPE') != \
                'CorrelationVolumeWarping' or _cfg_get(cfg, 'HEAD', 'TYPE') != 'DirectDeepResBlockMLP':
            raise NotImplementedError("only the FAR map-free recipe (ResUNet / CorrelationVolumeWarping / "
                                      "DirectDeepResBlockMLP) is on the path")
        self.encoder = ResUNet(_cfg_get(cfg, 'ENCODER'))
        self.max_steps, self.use_prior = max_steps, use_prior
        self.use_vanilla_transformer = use_vanilla_transformer
        if use_vanilla_transformer:
            self.transformer = nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model=256, nhead=8), num_layers=6,
                                                     enable_nested_tensor=False)
        self.aggregator = CorrelationVolumeWarping(_cfg_get(cfg, 'AGGREGATOR'), self.encoder.num_out_layers)
        self.head = DirectDeepResBlockMLP(cfg, self.aggregator.num_out_layers, full_forward_pass=False)
        self.H2, self.H, self.pose_size, self.num_corr_size = 512, 256 * 12 * 9, 9, 3
        self.use_loftr_preds = use_loftr_preds
        if use_loftr_preds:
            from .loftr import LoFTR, upstream_loftr_cfg
            from .solver import EssentialMatrixSolver
            self.matcher = LoFTR(matcher_config if matcher_config is not None else upstream_loftr_cfg()).eval()
            self.pose_solver = EssentialMatrixSolver(_cfg_get(cfg, 'SOLVER'), self.use_prior)
            self.moe_predictor = nn.Sequential(nn.Linear(self.H + 2 * self.pose_size + self.num_corr_size, self.H2),
                                               nn.ReLU(), nn.Linear(self.H2, self.H2), nn.ReLU(),
                                               nn.Linear(self.H2, 2), nn.Sigmoid())
        if use_vanilla_transformer:
            self.pose_regressor = nn.Sequential(nn.Linear(self.H, self.H2), nn.ReLU(), nn.Linear(self.H2, self.H2),
                                                nn.ReLU(), nn.Linear(self.H2, self.pose_size))
        if ckpt_path is not None:
            state_dict = torch.load(ckpt_path, map_location='cpu')['state_dict']
            self.load_state_dict(state_dict, strict=not not_strict)
        if use_loftr_preds:
            for prm in self.matcher.parameters():
                prm.requires_grad = False

    # ---- matcher + solver (model.py:167-172, 241-276), all B pairs at once ------------------------------------------
    @torch.no_grad()
    def match_batch(self, image0, image1):
        batch = {'image0': image0, 'image1': image1}
        self.matcher(batch)
        return batch

    @torch.no_grad()
    def solve_batch(self, matches, K0, K1, prior_rt=None, seed=0, inl_th=None):
        """The per-sample `pose_solver.estimate_pose` loop (model.py:245-273) as one batched GPU RANSAC round.
        Returns loftr_rt [B,3,4] (identity where no model, :268-271) and inliers [B,3] (use_prior) or [B,1]."""
        from .ransac import ransac_round
        B = K0.shape[0]
        dev = matches['mkpts0_f'].device
        K0, K1 = K0.to(dev).float(), K1.to(dev).float()
        prior = self.use_prior and prior_rt is not None
        th = 3e-7 if prior else (inl_th if inl_th is not None else self.pixel_threshold(K0, K1))
        r = ransac_round(matches['mkpts0_f'], matches['mkpts1_f'], matches['m_bids'], K0, K1,
                         prior_rt if prior else None, inl_th=th, seed=seed)
        ok = (r['best'] >= 0) & (r['n_pos'] > 0)
        rt = torch.where(ok[:, None, None], r['Rt'], torch.eye(3, 4, device=dev).expand(B, 3, 4))
        n = torch.where(ok, r['n_pos'], torch.zeros_like(r['n_pos'])).float()
        if self.use_prior:
            # model.py:257-262: [recoverPose inliers, tight, ultra tight]; the OpenCV (prior-free) round leaves the two
            # tight counters at 0 (pose_solver.py:45)
            c3 = r['counts3'].float() * ok[:, None]
            inl = torch.stack([n, c3[:, 1] if prior else torch.zeros_like(n), c3[:, 2] if prior else torch.zeros_like(n)], 1)
        else:
            inl = n[:, None]
        return rt, inl

    def pixel_threshold(self, K0, K1):
        """pose_solver.py:44: PIX_THRESHOLD / mean focal, squared (the round thresholds the SQUARED Sampson error).
        One scalar for the batch (pair 0's intrinsics; map-free batches share a camera model).  Reads K on the host:
        call it before any kernel of the forward is queued so the read does not wait on the stream."""
        k0, k1 = K0[0].detach().float().cpu(), K1[0].detach().float().cpu()
        f = float((k0[0, 0] + k1[1, 1] + k0[1, 1] + k1[0, 0]) / 4)
        return (float(self.pose_solver.ransac_pix_threshold) / f) ** 2

    def image_branch(self, data):
        """encoder x2 -> aggregator -> head (-> transformer): everything that does not depend on the solver."""
        vol0 = self.encoder(data['image0_reg'])
        vol1 = self.encoder(data['image1_reg'])
        global_volume = self.aggregator(vol0, vol1)
        R, t, feats = self.head(global_volume, data)
        if self.use_vanilla_transformer:
            B, C, H, W = feats.shape
            tokens = feats.reshape(B, C, H * W).permute(0, 2, 1).contiguous()          # [B, S, E] (reference: [S,B,E])
            feats = transformer_encoder(self.transformer, tokens).permute(0, 2, 1).contiguous()   # [B, C, S]
        return R, t, feats

    def transformer_head(self, features):
        B = features.shape[0]
        pr = self.pose_regressor
        pred = ops.linear(ops.linear(ops.linear(features.reshape(B, -1), pr[0].weight, pr[0].bias, ACT_RELU),
                                     pr[2].weight, pr[2].bias, ACT_RELU), pr[4].weight, pr[4].bias)
        return pred[..., 3:], pred[..., :3]

    def forward(self, data):
        priorRT = None
        num_loops = 2 if self.use_prior else 1
        matches = branch = None
        R = t = None
        inl_th = self.pixel_threshold(data['K_color0'], data['K_color1']) if self.use_loftr_preds else None
        for loop in range(num_loops):
            if self.use_loftr_preds:
                if matches is None:   # the matcher does not depend on the prior: evaluated once (reference: per loop)
                    matches = self.match_batch(data['image0'], data['image1'])
                    data['mkpts0_f'], data['mkpts1_f'], data['m_bids'] = matches['mkpts0_f'], matches['mkpts1_f'], \
                        matches['m_bids']
                data['loftr_rt'], data['inliers'] = self.solve_batch(matches, data['K_color0'], data['K_color1'],
                                                                     priorRT, seed=loop, inl_th=inl_th)
            with torch.no_grad():
                if branch is None:
                    branch = self.image_branch(data)
                R, t, feats = branch
                if self.use_loftr_preds:
                    R, t = self.regression_mlp(feats, data['loftr_rt'].float(), data['inliers'], R, t)
                    if self.use_prior and loop < num_loops - 1:
                        priorRT = torch.cat([rotation_6d_to_matrix(R), t.unsqueeze(2)], dim=-1)
                elif self.use_vanilla_transformer:
                    R, t = self.transformer_head(feats)
                else:
                    data['inliers'] = 0
        data['R'], data['t'] = R, t
        return R, t
