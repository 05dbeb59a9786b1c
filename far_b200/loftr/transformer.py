"""LoFTR linear-attention transformer and the FAR pose-fusion head, on the sm_100a kernels.

Same class names / constructor arguments / parameter names / forward signatures as
mp3d_loftr/src/loftr/loftr_module/transformer.py, so reference checkpoints load with strict=True and the classes
drop into LoFTR / PL_LoFTR unchanged.  Every forward() is one or a few C-ABI calls (far_b200.ops); nothing here
runs on the CPU or through PyTorch eager math beyond [N,9]-sized pose glue.

Batch semantics: the reference head only works for B=1.  Here a batch of N pairs means N
independent B=1 evaluations, vectorised (pair b = (feat0[b], feat1[b])).
"""
import copy
from functools import partial

import torch
import torch.nn as nn

from .. import ops
from .._lib import ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID, ENGINE_AUTO
from .pose import pose_mean_6d, pose_std_6d


class LinearAttention(nn.Module):
    """Q=elu(q)+1, K=elu(k)+1, out = Q (K^T V) / (Q . sum K)   (linear_attention.py:20-52)."""

    def __init__(self, eps=1e-6, use_num_corres=False):
        super().__init__()
        self.eps = eps

    def forward(self, queries, keys, values, q_mask=None, kv_mask=None, loftr_preds=None):
        if q_mask is not None or kv_mask is not None:
            raise NotImplementedError("padding masks (MegaDepth) are outside the FAR eval path (SURVEY.md 8a a2)")
        return ops.linear_attention(queries, keys, values, self.eps)


class LoFTREncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, attention='linear', use_num_corres=False):
        super().__init__()
        if attention != 'linear':
            raise NotImplementedError("only ATTENTION='linear' is used by the shipped configs (config/default.py:22,44,52)")
        self.dim = d_model // nhead
        self.nhead = nhead
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.attention = LinearAttention(use_num_corres=use_num_corres)
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.mlp = nn.Sequential(
            nn.Linear(d_model * 2, d_model * 2, bias=False),
            nn.ReLU(True),
            nn.Linear(d_model * 2, d_model, bias=False),
        )
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.engine = ENGINE_AUTO
        self.cache_weight_split = True

    def _weights(self):
        return {"q_proj": self.q_proj.weight, "k_proj": self.k_proj.weight, "v_proj": self.v_proj.weight,
                "merge": self.merge.weight, "mlp0": self.mlp[0].weight, "mlp2": self.mlp[2].weight,
                "norm1_w": self.norm1.weight, "norm1_b": self.norm1.bias,
                "norm2_w": self.norm2.weight, "norm2_b": self.norm2.bias}

    def _presplit(self, device):
        """tcgen05 operand splits of the five static weight blocks, cached on the module and refreshed when a weight's
        version counter / storage changes (load_state_dict, .to(), an optimizer step).  Inference only: under autograd
        or on the CPU nothing is cached."""
        ws = (self.q_proj.weight, self.k_proj.weight, self.v_proj.weight, self.merge.weight, self.mlp[0].weight,
              self.mlp[2].weight)
        key = tuple((w.data_ptr(), w._version) for w in ws) + (str(device),)
        c = getattr(self, "_far_presplit", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                kv = torch.cat([self.k_proj.weight, self.v_proj.weight], dim=0).contiguous()
                c = (key, {"q_proj": ops.tc_weight_split(self.q_proj.weight), "kv": ops.tc_weight_split(kv),
                           "merge": ops.tc_weight_split(self.merge.weight), "mlp0": ops.tc_weight_split(self.mlp[0].weight),
                           "mlp2": ops.tc_weight_split(self.mlp[2].weight)})
            self._far_presplit = c
        return c[1]

    def forward(self, x, source, x_mask=None, source_mask=None, loftr_preds=None):
        """x [N,L,C], source [N,S,C] -> [N,L,C]   (transformer.py:44-67)."""
        if x_mask is not None or source_mask is not None:
            raise NotImplementedError("padding masks are outside the FAR eval path")
        ps = self._presplit(x.device) if (x.is_cuda and self.cache_weight_split and not torch.is_grad_enabled()) else None
        return ops.loftr_encoder_layer(x, source, self._weights(), self.nhead, self.engine, self.norm1.eps,
                                       self.norm2.eps, ps)


class LocalFeatureTransformer(nn.Module):
    """Layer schedule of transformer.py:90-112: 'self' -> l(f0,f0), l(f1,f1); 'cross' -> f0=l(f0,f1), f1=l(f1,f0_new)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.d_model = config['d_model']
        self.nhead = config['nhead']
        self.layer_names = config['layer_names']
        if 'regress_use_num_corres' not in config:
            config['regress_use_num_corres'] = False
        layer = LoFTREncoderLayer(config['d_model'], config['nhead'], config['attention'], config['regress_use_num_corres'])
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(len(self.layer_names))])
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, feat0, feat1, mask0=None, mask1=None, loftr_preds=None, inv_loftr_preds=None):
        assert self.d_model == feat0.size(2), "the feature number of src and transformer must be equal"
        for layer, name in zip(self.layers, self.layer_names):
            if name == 'self':
                feat0 = layer(feat0, feat0, mask0, mask0)
                feat1 = layer(feat1, feat1, mask1, mask1)
            elif name == 'cross':
                feat0 = layer(feat0, feat1, mask0, mask1)
                feat1 = layer(feat1, feat0, mask1, mask0)
            else:
                raise KeyError(name)
        return feat0, feat1


def get_positional_encodings(B, N, intrinsics=None):
    """The 6-vector (y^2, x^2, xy, y, x, 1) per token with (x,y) = K^-1 [x_k, y_j, 1] on a linspace(-1,1) grid.
    The mp3d reference overwrites the intrinsics with constants and hard-codes h,w = 60,80
    (transformer.py:194-196), so the result is a CONSTANT [4800,6] table: computed once, vectorised, instead of
    the reference's 4800-iteration python loop per forward (:236-240)."""
    h, w = 60, 80
    assert N == h * w, "the FAR-LoFTR head is hard-wired to a 60x80 coarse grid (640x480 input)"
    fx, fy, cx, cy = 517 / 9, 517 / 8, 40.0, 30.0
    ys = torch.linspace(-1, 1, steps=h)
    xs = torch.linspace(-1, 1, steps=w)
    fxn, fyn = torch.tensor(fx / (cx * 2) * 2), torch.tensor(fy / (cy * 2) * 2)
    cxn, cyn = torch.tensor(cx / (cx * 2) * 2 - 1), torch.tensor(cy / (cy * 2) * 2 - 1)
    K = torch.zeros(3, 3)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = fxn, fyn, cxn, cyn, 1.0
    Kinv = torch.inverse(K)
    jj, kk = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    pts = torch.stack([xs[kk.reshape(-1)], ys[jj.reshape(-1)], torch.ones(h * w)], 0)
    wv = Kinv @ pts
    p4, p3 = wv[0] / wv[2], wv[1] / wv[2]
    table = torch.stack([p3 * p3, p4 * p4, p3 * p4, p3, p4, torch.ones(h * w)], dim=1)
    return table.unsqueeze(0).expand(B, N, 6)


class CrossAttention(nn.Module):
    """Dual-softmax bilinear "essential-matrix module" attention (transformer.py:250-303)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj_fundamental = nn.Linear(dim + int(6 * self.num_heads), dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self._pos = None
        self.engine = ENGINE_AUTO

    def positional(self, N, device):
        if self._pos is None or self._pos.device != device or self._pos.shape[1] != N:
            self._pos = get_positional_encodings(1, N).contiguous().to(device)
        return self._pos

    def forward(self, x1, x2, intrinsics=None, loftr_preds=None, inv_loftr_preds=None):
        B, N, C = x1.shape
        h = self.num_heads
        qkv1 = ops.linear(x1, self.qkv.weight, self.qkv.bias)  # [B,N,3C] == [B,N,3,h,d]
        qkv2 = ops.linear(x2, self.qkv.weight, self.qkv.bias)
        f1, f2 = ops.emm_bilinear_attn(qkv1, qkv2, self.positional(N, x1.device), h, self.scale, self.engine)
        ch = C + 6 * h
        # [B,h,d+6,d+6] -> reshape(B, C+6h, (C+6h)/h).transpose(-2,-1) (:294-295): 70x280 per pair, a view + tiny copy
        f1 = f1.reshape(B, ch, ch // h).transpose(-2, -1)
        f2 = f2.reshape(B, ch, ch // h).transpose(-2, -1)
        f2 = ops.linear(f2, self.proj_fundamental.weight, self.proj_fundamental.bias)
        f1 = ops.linear(f1, self.proj_fundamental.weight, self.proj_fundamental.bias)
        return f2, f1  # flipped, as in the reference (:300-303)


class Mlp(nn.Module):
    """timm-style MLP (vit_layers/mlp.py): fc1 -> GELU(erf) -> fc2."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return ops.linear(ops.linear(x, self.fc1.weight, self.fc1.bias, ACT_GELU), self.fc2.weight, self.fc2.bias)


class CrossBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, use_pos_embedding=False, distilled=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.cross_attn = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.patch_embed = nn.Identity()
        self.h, self.w = 60, 80
        self.num_tokens = 0
        self.pos_embed = 0
        if use_pos_embedding:
            self.pos_embed = nn.Parameter(torch.zeros(1, self.h * self.w + self.num_tokens, 256))
            nn.init.trunc_normal_(self.pos_embed, std=.02)
        for m in self.modules():  # _init_vit_weights (transformer.py:150-181), default branch
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward_pairs(self, feat0, feat1):
        """feat0, feat1 [N,hw,C] -> fundamental [N, 2*(d+6), C]; pair b = (feat0[b], feat1[b])."""
        pe = self.pos_embed[0] if isinstance(self.pos_embed, torch.Tensor) else None
        n1 = ops.layernorm(feat0, self.norm1.weight, self.norm1.bias, self.norm1.eps, pre_add=pe)
        n2 = ops.layernorm(feat1, self.norm1.weight, self.norm1.bias, self.norm1.eps, pre_add=pe)
        f1, f2 = self.cross_attn(n1, n2)
        fund = torch.cat([f1, f2], dim=1)  # [N, 2*(d+6), C]: the (pair, image) interleave of :345-346
        return fund + self.mlp(ops.layernorm(fund, self.norm2.weight, self.norm2.bias, self.norm2.eps))

    def forward(self, x, intrinsics=None, loftr_preds=None, inv_loftr_preds=None):
        """Reference layout: x = cat([feat0, feat1], 0) for ONE pair -> [2, d+6, C] (transformer.py:335-348)."""
        b_s, h_w, nf = x.shape
        if b_s != 2:
            raise ValueError("CrossBlock.forward keeps the reference's B=1 layout; use forward_pairs for batches")
        out = self.forward_pairs(x[0:1], x[1:2])
        return out.reshape(b_s, -1, nf)


def _seq_linear(seq, x, acts):
    """nn.Sequential of Linear/activation pairs through the CUDA GEMM (activations fused into the epilogue)."""
    lin = [m for m in seq if isinstance(m, nn.Linear)]
    for m, a in zip(lin, acts):
        x = ops.linear(x, m.weight, m.bias, a)
    return x


class LocalFeatureTransformerRegressor(nn.Module):
    """LoFTR layers + EMM head (transformer.py:350-499)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        num_heads, feat_size, pos_enc = 4, 256, 6
        pose_size_in = pose_size = 9
        self.pose_size = pose_size
        if self.config['regress']['regress_use_num_corres']:
            pose_size_in += 1
        if self.config['use_many_ransac_thr']:
            pose_size_in += 3
        self.pose_size_in = pose_size_in
        self.H = int(num_heads * 2 * (feat_size // num_heads + pos_enc) * (feat_size // num_heads))
        self.H2 = 512
        rc = config['regress']
        if rc['use_simple_moe']:
            self.encoder = nn.Sequential(nn.Linear(self.H, self.H2), nn.ReLU(), nn.Linear(self.H2, self.H2))
            local = 1 if rc['use_1wt'] else (2 if rc['use_2wt'] else pose_size)
            self.moe_predictor = nn.Sequential(
                nn.Linear(self.H + pose_size + pose_size_in, self.H2), nn.ReLU(),
                nn.Linear(self.H2, self.H2), nn.ReLU(),
                nn.Linear(self.H2, local), nn.Sigmoid())
            self.pose_regressor_simple_moe = nn.Sequential(nn.Linear(self.H2, self.H2), nn.ReLU(),
                                                           nn.Linear(self.H2, pose_size))
        else:
            self.pose_regressor = nn.Sequential(nn.Linear(self.H, self.H2), nn.ReLU(), nn.Linear(self.H2, self.H2),
                                                nn.ReLU(), nn.Linear(self.H2, pose_size))
        self.norm = partial(nn.LayerNorm, eps=1e-6)(feat_size)
        self.emm = CrossBlock(dim=feat_size, num_heads=num_heads, qkv_bias=True,
                              use_pos_embedding=rc['use_pos_embedding'])
        if config['regress_loftr_layers'] > 0:
            self.loftr = LocalFeatureTransformer(config['regress'])

    def forward_trunk(self, feat0, feat1, run_loftr=True):
        """The part of forward()/forward_emm() that is a pure function of the two feature maps (transformer.py:423-432,
        448-456): regress LoFTR layers -> EMM CrossBlock -> outer LayerNorm -> flatten [B,35840]; with use_simple_moe
        also `encoder` and `pose_regressor_simple_moe` (the regressed 9-D pose).  The solver prediction only enters
        afterwards (forward_gate), so a caller that invokes the head twice on the same feature maps
        (fine_pred_steps = 2, lightning_loftr.py:338-346) can evaluate this once per forward."""
        if run_loftr and self.config['regress_loftr_layers'] > 0:
            feat0, feat1 = self.loftr(feat0, feat1)
        B = feat0.shape[0]
        x = self.emm.forward_pairs(feat0, feat1)                                   # [B, 140, 256]
        features = ops.layernorm(x, self.norm.weight, self.norm.bias, self.norm.eps).reshape(B, -1)  # [B, 35840]
        pred_reg_6d = None
        if self.config['regress']['use_simple_moe']:
            feats = _seq_linear(self.encoder, features, (ACT_RELU, ACT_NONE))
            pred_reg_6d = _seq_linear(self.pose_regressor_simple_moe, feats, (ACT_RELU, ACT_NONE))
        return features, pred_reg_6d

    def forward_gate(self, trunk, loftr_preds=None, inv_loftr_preds=None):
        """Everything that depends on the solver prediction (transformer.py:457-469): gate MLP on
        cat([features, regressed, solver]) and the gated blend."""
        features, pred_reg_6d = trunk
        rc = self.config['regress']
        if not rc['use_simple_moe']:
            return _seq_linear(self.pose_regressor, features, (ACT_RELU, ACT_RELU, ACT_NONE)), \
                (features if rc['save_mlp_feats'] else None), None
        tail = torch.cat([pred_reg_6d, loftr_preds.float()], dim=-1)              # [B, 9 + pose_size_in]
        m0 = self.moe_predictor[0]
        # moe_predictor.0 on cat([features, pred, solver]): two K-segments, no [B,35862] concat (transformer.py:458-459)
        hid = ops.linear_cat_tail(features, tail, m0, ACT_RELU)
        hid = ops.linear(hid, self.moe_predictor[2].weight, self.moe_predictor[2].bias, ACT_RELU)
        pred_RT_wt = ops.linear(hid, self.moe_predictor[4].weight, self.moe_predictor[4].bias, ACT_SIGMOID)
        if rc['use_2wt'] and not rc['use_5050_weight']:
            wt2 = pred_RT_wt
        elif rc['use_2wt']:
            wt2 = torch.full_like(pred_RT_wt, 0.5)
        else:
            wt2 = pred_RT_wt[..., :1].expand(-1, 2).contiguous()
        dev = features.device
        pose_preds = ops.pose_blend_mp3d(pred_reg_6d, loftr_preds.float(), wt2, pose_mean_6d.to(dev),
                                         pose_std_6d.to(dev), rc['scale_8pt'])
        return pose_preds, (features if rc['save_mlp_feats'] else None), pred_RT_wt

    def forward_emm(self, feat0, feat1, loftr_preds=None, inv_loftr_preds=None):
        """forward_emm of the reference takes the post-`self.loftr` features (transformer.py:448)."""
        return self.forward_gate(self.forward_trunk(feat0, feat1, run_loftr=False), loftr_preds, inv_loftr_preds)

    def forward(self, feat0, feat1, loftr_preds=None, inv_loftr_preds=None, mask0=None, mask1=None, F=None):
        return self.forward_gate(self.forward_trunk(feat0, feat1), loftr_preds, inv_loftr_preds)
