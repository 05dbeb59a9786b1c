"""8pt-ViT (`ViTEss`) with the reference's interface (interiornetStreetlearn_8ptVit/src/model.py:38-217):
`ViTEss(args, global_pose_mean, global_pose_std)`, `forward(images, intrinsics=None, inference=False,
loftr_num_corr=None, loftr_preds=None)` -> (t [B,3], rot [B,J,3], R [B,3,3], r6d [B,6]).

Feature extraction (ResNet-18 stem + ResidualBlock) stays on cuDNN; the 5 ViT blocks, the dual-softmax bilinear
CrossBlock and the gated pose MLPs run on the sm_100a kernels (far_b200.ops).  Parameter names match the
reference checkpoint layout (`resnet.*`, `extractor_final_conv.*`, `fusion_transformer.{pos_embed,blocks.N.*,norm}`,
`pose_regressor.*`, `moe_predictor.*`)."""
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID


class ResidualBlock(nn.Module):
    """src/modules/extractor.py:5-65 ('batch' norm, kernel_size > 1 variant used by ViTEss)."""

    def __init__(self, in_planes, planes, norm_fn='batch', stride=1, kernel_size=1):
        super().__init__()
        assert norm_fn == 'batch'
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=kernel_size) if kernel_size > 1 else \
            nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = nn.BatchNorm2d(planes)
        self.norm2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride > 1 or kernel_size > 1:
            self.norm3 = nn.BatchNorm2d(planes)
            k = 1 if stride > 1 else kernel_size
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=k, stride=stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


def get_positional_encodings(B, N, intrinsics=None):
    """(y^2, x^2, xy, y, x, 1) per token (src/modules/vision_transformer.py:90-158), vectorised.  Keeps the
    reference's index quirk: the value computed from (xs[k], ys[j]) lands at token k*w + j (:150-151)."""
    h, w = (24, 24) if N == 24 * 24 else (48, 64)
    assert N == h * w, 'unexpected resolution for positional encoding'
    ys = torch.linspace(-1, 1, steps=h)
    xs = torch.linspace(-1, 1, steps=w)
    if intrinsics is None:
        p3 = ys.unsqueeze(0).repeat(B, w)
        p4 = xs.repeat_interleave(h).unsqueeze(0).repeat(B, 1)
    else:
        intr = intrinsics.detach().float().cpu()
        assert torch.all(intr[:, 0] == intr[:, 1])
        fx, fy, cx, cy = intr[:, 0].unbind(dim=-1)
        K = torch.zeros(B, 3, 3)
        K[:, 0, 0] = (fx / (cx * 2)) * 2
        K[:, 1, 1] = (fy / (cy * 2)) * 2
        K[:, 0, 2] = (cx / (cx * 2)) * 2 - 1
        K[:, 1, 2] = (cy / (cy * 2)) * 2 - 1
        K[:, 2, 2] = 1
        Kinv = torch.inverse(K)
        kk, jj = torch.meshgrid(torch.arange(w), torch.arange(h), indexing='ij')   # token index = k*w + j
        idx = (kk * w + jj).reshape(-1)
        pts = torch.stack([xs[kk.reshape(-1)], ys[jj.reshape(-1)], torch.ones(h * w)], 0)  # [3, hw]
        wv = Kinv @ pts                                                                     # [B, 3, hw]
        p3 = ys.unsqueeze(0).repeat(B, w).clone()
        p4 = xs.repeat_interleave(h).unsqueeze(0).repeat(B, 1).clone()
        p3[:, idx] = wv[:, 1] / wv[:, 2]
        p4[:, idx] = wv[:, 0] / wv[:, 2]
    return torch.stack([p3 * p3, p4 * p4, p3 * p4, p3, p4, torch.ones(B, N)], dim=2)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return ops.linear(ops.linear(x, self.fc1.weight, self.fc1.bias, ACT_GELU), self.fc2.weight, self.fc2.bias)


class Attention(nn.Module):
    """timm attention (vision_transformer.py:236-262): qkv Linear -> softmax(q k^T * scale) v -> proj."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.attn_drop = nn.Dropout(attn_drop)

    def forward(self, x, camera=None, intrinsics=None):
        qkv = ops.linear(x, self.qkv.weight, self.qkv.bias)
        return ops.linear(ops.softmax_attention(qkv, self.num_heads, self.scale), self.proj.weight, self.proj.bias)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, camera=None, intrinsics=None):
        x = x + self.attn(ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps))
        return x + self.mlp(ops.layernorm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps))


class CrossAttention(nn.Module):
    """vision_transformer.py:160-208."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj_fundamental = nn.Linear(dim + int(6 * self.num_heads), dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x1, x2, camera=None, intrinsics=None):
        B, N, C = x1.shape
        h = self.num_heads
        pos = get_positional_encodings(B, N, intrinsics=intrinsics).to(x1.device)
        qkv1 = ops.linear(x1, self.qkv.weight, self.qkv.bias)
        qkv2 = ops.linear(x2, self.qkv.weight, self.qkv.bias)
        f1, f2 = ops.emm_bilinear_attn(qkv1, qkv2, pos, h, self.scale)
        ch = C + 6 * h
        f1 = f1.reshape(B, ch, ch // h).transpose(-2, -1)
        f2 = f2.reshape(B, ch, ch // h).transpose(-2, -1)
        f2 = ops.linear(f2, self.proj_fundamental.weight, self.proj_fundamental.bias)
        f1 = ops.linear(f1, self.proj_fundamental.weight, self.proj_fundamental.bias)
        return f2, f1


class CrossBlock(nn.Module):
    """vision_transformer.py:210-234: x [2B,N,C] holds the pairs interleaved (reshape(-1, 2, N, C))."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.cross_attn = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, camera=None, intrinsics=None):
        b_s, h_w, nf = x.shape
        x = x.reshape(-1, 2, h_w, nf)
        n1 = ops.layernorm(x[:, 0], self.norm1.weight, self.norm1.bias, self.norm1.eps)
        n2 = ops.layernorm(x[:, 1], self.norm1.weight, self.norm1.bias, self.norm1.eps)
        f1, f2 = self.cross_attn(n1, n2, camera, intrinsics=intrinsics)
        fund = torch.cat([f1.unsqueeze(1), f2.unsqueeze(1)], dim=1).reshape(b_s, -1, nf)
        return fund + self.mlp(ops.layernorm(fund, self.norm2.weight, self.norm2.bias, self.norm2.eps))


class VisionTransformer(nn.Module):
    """The part of timm's ViT that ViTEss keeps (vision_transformer.py:286-392): pos_embed, `depth` blocks whose last
    one is the CrossBlock, final LayerNorm(eps=1e-6).  patch_embed/head are Identity and cls_token is None there."""

    def __init__(self, num_patches, embed_dim=192, depth=6, num_heads=3, mlp_ratio=4., qkv_bias=True):
        super().__init__()
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        nn.init.xavier_uniform_(self.pos_embed)
        self.pos_drop = nn.Dropout(0.)
        blocks = [Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer=norm_layer) for _ in range(depth - 1)]
        blocks.append(CrossBlock(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer=norm_layer))
        self.blocks = nn.Sequential(*blocks)
        self.norm = norm_layer(embed_dim)
        self.patch_embed = nn.Identity()
        self.head = nn.Identity()
        self.cls_token = None
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)


def compute_rotation_matrix_from_ortho6d(ortho6d):
    """src/geom/RotationContinuity/sanity_test/code/tools.py:47-62 (columns x, y, z)."""
    x_raw, y_raw = ortho6d[:, 0:3], ortho6d[:, 3:6]

    def nrm(v):
        return v / torch.clamp(torch.sqrt(v.pow(2).sum(1, keepdim=True)), min=1e-8)

    x = nrm(x_raw)
    z = nrm(torch.cross(x, y_raw, dim=1))
    y = torch.cross(z, x, dim=1)
    return torch.stack((x, y, z), dim=2)


def compute_pose_from_rotation_matrix(T_pose, r_matrix):
    """tools.py:9-17."""
    return torch.matmul(r_matrix[:, None], T_pose.to(r_matrix)[None, :, :, None])[..., 0]


class ViTEss(nn.Module):
    def __init__(self, args, global_pose_mean=None, global_pose_std=None):
        super().__init__()
        import torchvision.models as models
        self.total_num_features = 192
        self.feature_resolution = (24, 24)
        self.pose_size = 9
        self.num_patches = 24 * 24
        self.H2 = args.fc_hidden_size
        self.use_loftr_gating = args.use_loftr_gating
        self.global_pose_mean, self.global_pose_std = global_pose_mean, global_pose_std
        self.use_normalized_6d = args.use_normalized_6d
        self.flatten = nn.Flatten(0, 1)
        self.resnet = models.resnet18(weights=None)  # the reference downloads ImageNet weights here (model.py:58)
        self.resnet.fc = nn.Identity()
        self.extractor_final_conv = ResidualBlock(128, self.total_num_features, 'batch', kernel_size=max(1, 28 - 24 + 1))
        if not args.fusion_transformer:
            raise NotImplementedError("only the fusion_transformer variant is a FAR recipe")
        self.num_heads = 3
        self.transformer_depth = args.transformer_depth
        self.fusion_transformer = VisionTransformer(self.num_patches, self.total_num_features, args.transformer_depth,
                                                    self.num_heads)
        self.H = int(self.num_heads * 2 * (self.total_num_features // self.num_heads + 6) *
                     (self.total_num_features // self.num_heads))
        if self.use_loftr_gating:
            self.moe_predictor = nn.Sequential(nn.Linear(self.H + 2 * self.pose_size + 1, self.H2), nn.ReLU(),
                                               nn.Linear(self.H2, self.H2), nn.ReLU(), nn.Linear(self.H2, 2), nn.Sigmoid())
        self.pose_regressor = nn.Sequential(nn.Linear(self.H, self.H2), nn.ReLU(), nn.Linear(self.H2, self.H2), nn.ReLU(),
                                            nn.Linear(self.H2, self.pose_size))
        self.T_pose = args.T_pose

    def update_intrinsics(self, input_shape, intrinsics):
        sizey, sizex = self.feature_resolution
        intrinsics[:, :, [0, 2]] = (sizex / input_shape[-1]) * intrinsics[:, :, [0, 2]]   # mutates the caller's tensor,
        intrinsics[:, :, [1, 3]] = (sizey / input_shape[-2]) * intrinsics[:, :, [1, 3]]   # like model.py:126-127
        return intrinsics

    def extract_features(self, images, intrinsics=None):
        if images.is_cuda and not self.training and not torch.is_grad_enabled() and images.dtype == torch.float32 \
                and not getattr(self, 'force_eager', False):
            return self._extract_features_fused(images, intrinsics)
        images = images[:, :, [2, 1, 0]] / 255.0
        mean = torch.as_tensor([0.485, 0.456, 0.406], device=images.device)
        std = torch.as_tensor([0.229, 0.224, 0.225], device=images.device)
        images = images.sub_(mean[:, None, None]).div_(std[:, None, None])
        if intrinsics is not None:
            intrinsics = self.update_intrinsics(images.shape, intrinsics)
        x = F.interpolate(self.flatten(images), size=224)
        r = self.resnet
        x = r.layer2(r.layer1(r.maxpool(r.relu(r.bn1(r.conv1(x))))))
        x = self.extractor_final_conv(x)
        n = x.shape[0]
        feats = x.reshape(n, -1, self.num_patches)[:, :self.total_num_features].permute(0, 2, 1)
        return feats, intrinsics

    # ---- eval-time fused path (same arithmetic): one preprocessing kernel, eval BatchNorm folded into the convolutions,
    # cuDNN conv+bias(+residual)+ReLU fused calls (channels_last) -- removes ~45 elementwise / BN launches per batch
    def _folded(self):
        mods = [self.resnet.conv1, self.resnet.bn1, self.resnet.layer1, self.resnet.layer2, self.extractor_final_conv]
        ts = [t for m in mods for t in list(m.parameters()) + list(m.buffers())]
        key = tuple(int(t._version) for t in ts) + (self.resnet.conv1.weight.data_ptr(),)
        c = getattr(self, "_fold_cache", None)
        if c is not None and c[0] == key:
            return c[1]

        def fold(conv, bn):
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = (conv.weight * scale[:, None, None, None]).contiguous(memory_format=torch.channels_last)
            b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
            return w, ((b0 - bn.running_mean) * scale + bn.bias).contiguous()

        with torch.no_grad():
            f = {"stem": fold(self.resnet.conv1, self.resnet.bn1)}
            for li, layer in enumerate((self.resnet.layer1, self.resnet.layer2)):
                for bi, blk in enumerate(layer):
                    f[(li, bi, 1)] = fold(blk.conv1, blk.bn1)
                    f[(li, bi, 2)] = fold(blk.conv2, blk.bn2)
                    if blk.downsample is not None:
                        wd, bd = fold(blk.downsample[0], blk.downsample[1])
                        f[(li, bi, "d")] = wd
                        f[(li, bi, 2)] = (f[(li, bi, 2)][0], (f[(li, bi, 2)][1] + bd).contiguous())
            e = self.extractor_final_conv
            f["e1"], f["e2"] = fold(e.conv1, e.norm1), fold(e.conv2, e.norm2)
            if e.downsample is not None:
                f["ed"] = fold(e.downsample[0], e.downsample[1])
        self._fold_cache = (key, f)
        return f

    def _extract_features_fused(self, images, intrinsics):
        B = images.shape[0]
        shape = images.shape
        x = ops.vit_preprocess(images.reshape(B * 2, 3, shape[-2], shape[-1]).contiguous(), 224)
        if intrinsics is not None:
            intrinsics = self.update_intrinsics(shape, intrinsics)
        f = self._folded()
        one, pad1 = (1, 1), (1, 1)
        r = self.resnet
        w, b = f["stem"]
        x = torch.relu_(F.conv2d(x.contiguous(memory_format=torch.channels_last), w, b, stride=2, padding=3))
        x = r.maxpool(x)
        for li, layer in enumerate((r.layer1, r.layer2)):
            for bi, blk in enumerate(layer):
                w1, b1 = f[(li, bi, 1)]
                w2, b2 = f[(li, bi, 2)]
                y = torch.cudnn_convolution_relu(x, w1, b1, blk.conv1.stride, pad1, one, 1)
                if blk.downsample is not None:   # relu(conv2(y) + b2 + conv_d(x) + b_d): the bias rides on conv2's
                    x = F.conv2d(x, f[(li, bi, "d")], None, stride=blk.downsample[0].stride)
                x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, one, pad1, one, 1)
        e = self.extractor_final_conv
        y = torch.cudnn_convolution_relu(x, f["e1"][0], f["e1"][1], e.conv1.stride, e.conv1.padding, one, 1)
        y = torch.relu_(F.conv2d(y, f["e2"][0], f["e2"][1], e.conv2.stride, e.conv2.padding))
        if e.downsample is not None:
            x = F.conv2d(x, f["ed"][0], f["ed"][1], e.downsample[0].stride, e.downsample[0].padding)
        x = torch.relu_(x + y)
        n = x.shape[0]
        feats = x.contiguous().reshape(n, -1, self.num_patches)[:, :self.total_num_features].permute(0, 2, 1)
        return feats, intrinsics

    def fusion_head(self, features, intrinsics, loftr_num_corr, loftr_preds):
        """Everything after extract_features (model.py:170-217), on the CUDA kernels.  features [2B,576,192]."""
        B = features.shape[0] // 2
        ft = self.fusion_transformer
        x = features.contiguous() + ft.pos_embed
        for layer in range(self.transformer_depth):
            x = ft.blocks[layer](x, intrinsics=intrinsics)
        feats = ops.layernorm(x, ft.norm.weight, ft.norm.bias, ft.norm.eps).reshape(B, -1)
        pr = self.pose_regressor
        pred = ops.linear(ops.linear(ops.linear(feats, pr[0].weight, pr[0].bias, ACT_RELU), pr[2].weight, pr[2].bias,
                                     ACT_RELU), pr[4].weight, pr[4].bias)
        if not self.use_loftr_gating:
            return pred, None
        dev = feats.device
        lp = loftr_preds.float().to(dev)
        l9 = torch.cat([lp[..., :3, 3], lp[..., :2, :3].reshape(B, 6)], dim=-1)
        if self.use_normalized_6d:
            l9 = (l9 - self.global_pose_mean.to(dev)) / self.global_pose_std.to(dev)
        l10 = torch.cat([l9, loftr_num_corr.detach().float().to(dev).unsqueeze(1) / 500], dim=-1)
        mp = self.moe_predictor
        tail = torch.cat([pred, l10], dim=-1)
        hid = ops.linear_cat_tail(feats, tail, mp[0], ACT_RELU)
        wt = ops.linear(ops.linear(hid, mp[2].weight, mp[2].bias, ACT_RELU), mp[4].weight, mp[4].bias, ACT_SIGMOID)
        pred_T = wt[..., :1] * pred[..., :3] + (1 - wt[..., :1]) * l10[..., :3]
        pred_R = wt[..., 1:] * pred[..., 3:] + (1 - wt[..., 1:]) * l10[..., 3:-1]
        return torch.cat([pred_T, pred_R], dim=-1), wt

    def forward(self, images, intrinsics=None, inference=False, loftr_num_corr=None, loftr_preds=None):
        features, intrinsics = self.extract_features(images, intrinsics)
        pose_preds, _ = self.fusion_head(features, intrinsics, loftr_num_corr, loftr_preds)
        rot6, tran = pose_preds[:, 3:], pose_preds[:, :3]
        if self.use_normalized_6d:
            dev = pose_preds.device
            r6u = rot6 * self.global_pose_std[3:].to(dev) + self.global_pose_mean[3:].to(dev)
            tu = tran * self.global_pose_std[:3].to(dev) + self.global_pose_mean[:3].to(dev)
        else:
            r6u, tu = rot6, tran
        R = compute_rotation_matrix_from_ortho6d(r6u)
        return tu, compute_pose_from_rotation_matrix(self.T_pose, R), R, rot6
