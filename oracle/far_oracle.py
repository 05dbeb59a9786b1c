"""CPU oracle for the FAR per-pair pose hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module, and only as the checker / the CPU
baseline.  `far_b200/` never imports it and has no CPU fallback.

What it is: a plain restatement, in torch-CPU fp32 tensor ops, of the reference's algorithm for the
path.  The reference (crockwell/far) is itself pure PyTorch, so "restating" means: same arithmetic, same
operation order, written functionally over a flat `state_dict` (reference parameter names) instead
of nn.Modules, with no dependency on the reference tree, yacs, kornia, lightning or OpenCV.
Every function cites the reference file:line it follows (paths relative to /root/reference).

Parity pin: `tests/golden/make_golden.py` runs the UNMODIFIED reference (imported through
`oracle/ref_import.py`) and this restatement on identical seeded inputs/weights and asserts agreement,
then commits the reference's outputs as fixtures under `tests/golden/`; `tests/test_oracle_golden.py`
re-checks the restatement against those fixtures on every run (no /root/reference needed).
Unpinned call sites (third-party arithmetic that is not under /root/reference): the two kornia 0.7.1
functions used by FineMatching (restated from kornia's public definition) and OpenCV's
findEssentialMat/recoverPose (not restated - the in-repo run_8point / decompose_essential_matrix are the
solver oracle).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# mp3d pose normalisation constants: mp3d_loftr/src/losses/loftr_loss.py:7-8
POSE_MEAN_6D = torch.tensor([-0.34898765, 0.17085525, -0.87944315, 0.50275223, 0.03533648, -0.18179045,
                             -0.03533648, 0.98189617, 0.09313615])
POSE_STD_6D = torch.tensor([1.94014405, 0.36770130, 1.88317520, 0.51837117, 0.12717603, 0.65426397,
                            0.12717603, 0.0188729, 0.09709263])


def _sub(sd, prefix):
    """state-dict view below `prefix.`"""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


# --------------------------------------------------------------------------------------- a1
def position_encoding_sine(d_model, h, w, temp_bug_fix=True):
    """mp3d_loftr/src/loftr/utils/position_encoding.py:22-35 (table), returned as [d_model, h, w]."""
    pe = torch.zeros((d_model, h, w))
    y_position = torch.ones((h, w)).cumsum(0).float().unsqueeze(0)
    x_position = torch.ones((h, w)).cumsum(1).float().unsqueeze(0)
    if temp_bug_fix:
        div_term = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / (d_model // 2)))
    else:  # position_encoding.py:28 -- `-log(1e4) / d_model // 2` floor-divides
        div_term = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / d_model // 2))
    div_term = div_term[:, None, None]
    pe[0::4] = torch.sin(x_position * div_term)
    pe[1::4] = torch.cos(x_position * div_term)
    pe[2::4] = torch.sin(y_position * div_term)
    pe[3::4] = torch.cos(y_position * div_term)
    return pe


def add_pos_and_flatten(feat_nchw, temp_bug_fix=True):
    """position_encoding.py:37-42 + rearrange 'n c h w -> n (h w) c' (loftr.py:100-101)."""
    n, c, h, w = feat_nchw.shape
    x = feat_nchw + position_encoding_sine(c, h, w, temp_bug_fix)[None]
    return x.permute(0, 2, 3, 1).reshape(n, h * w, c)


# --------------------------------------------------------------------------------------- a3
def linear_attention(q, k, v, eps=1e-6):
    """mp3d_loftr/src/loftr/loftr_module/linear_attention.py:20-52.  q [N,L,H,D]; k,v [N,S,H,D]."""
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    S = v.size(1)
    v = v / S
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + eps)
    return (torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * S).contiguous()


# --------------------------------------------------------------------------------------- a2
def loftr_encoder_layer(p, x, source, nhead):
    """mp3d_loftr/src/loftr/loftr_module/transformer.py:44-67 (masks None)."""
    bs, _, c = x.shape
    dim = c // nhead
    q = F.linear(x, p["q_proj.weight"]).view(bs, -1, nhead, dim)
    k = F.linear(source, p["k_proj.weight"]).view(bs, -1, nhead, dim)
    v = F.linear(source, p["v_proj.weight"]).view(bs, -1, nhead, dim)
    msg = linear_attention(q, k, v)
    msg = F.linear(msg.view(bs, -1, nhead * dim), p["merge.weight"])
    msg = F.layer_norm(msg, (c,), p["norm1.weight"], p["norm1.bias"], 1e-5)
    msg = F.linear(torch.cat([x, msg], dim=2), p["mlp.0.weight"])
    msg = F.linear(F.relu(msg), p["mlp.2.weight"])
    msg = F.layer_norm(msg, (c,), p["norm2.weight"], p["norm2.bias"], 1e-5)
    return x + msg


# --------------------------------------------------------------------------------------- a4
def local_feature_transformer(p, feat0, feat1, layer_names, nhead):
    """transformer.py:90-112.  NB the second cross call sees the UPDATED feat0 (:107-108)."""
    for i, name in enumerate(layer_names):
        lp = _sub(p, f"layers.{i}")
        if name == "self":
            feat0 = loftr_encoder_layer(lp, feat0, feat0, nhead)
            feat1 = loftr_encoder_layer(lp, feat1, feat1, nhead)
        elif name == "cross":
            feat0 = loftr_encoder_layer(lp, feat0, feat1, nhead)
            feat1 = loftr_encoder_layer(lp, feat1, feat0, nhead)
        else:
            raise KeyError(name)
    return feat0, feat1


# --------------------------------------------------------------------------------------- a5/a6
def dual_softmax_conf(feat_c0, feat_c1, temperature=0.1):
    """mp3d_loftr/src/loftr/utils/coarse_matching.py:105-118 (dual_softmax branch, no masks)."""
    c = feat_c0.shape[-1]
    f0 = feat_c0 / c ** .5
    f1 = feat_c1 / c ** .5
    sim = torch.einsum("nlc,nsc->nls", f0, f1) / temperature
    return F.softmax(sim, 1) * F.softmax(sim, 2)


def coarse_match_from_conf(conf, hw0_c, hw1_c, thr, border_rm, scale=8.0):
    """coarse_matching.py:149-265, eval branch (no `self.training` sampling, no scale0/scale1, no mask0).

    Returns dict with b_ids,i_ids,j_ids (int64, ascending (b,i)), mconf, mkpts0_c, mkpts1_c, m_bids."""
    n = conf.shape[0]
    h0, w0 = hw0_c
    h1, w1 = hw1_c
    mask = conf > thr  # strict > (:174)
    mask = mask.view(n, h0, w0, h1, w1).clone()
    b = border_rm
    if b > 0:  # mask_border, :8-25
        mask[:, :b] = False
        mask[:, :, :b] = False
        mask[:, :, :, :b] = False
        mask[:, :, :, :, :b] = False
        mask[:, -b:] = False
        mask[:, :, -b:] = False
        mask[:, :, :, -b:] = False
        mask[:, :, :, :, -b:] = False
    mask = mask.view(n, h0 * w0, h1 * w1)
    mask = mask * (conf == conf.max(dim=2, keepdim=True)[0]) * (conf == conf.max(dim=1, keepdim=True)[0])
    mask_v, all_j_ids = mask.max(dim=2)  # first True per row (:192)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j_ids[b_ids, i_ids]
    mconf = conf[b_ids, i_ids, j_ids]
    mk0 = torch.stack([i_ids % w0, torch.div(i_ids, w0, rounding_mode="floor")], dim=1) * scale
    mk1 = torch.stack([j_ids % w1, torch.div(j_ids, w1, rounding_mode="floor")], dim=1) * scale
    keep = mconf != 0
    return {"b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids, "gt_mask": mconf == 0, "m_bids": b_ids[keep],
            "mkpts0_c": mk0[keep], "mkpts1_c": mk1[keep], "mconf": mconf[keep]}


def coarse_matching(feat_c0, feat_c1, hw0_c, hw1_c, thr=0.2, border_rm=2, temperature=0.1, scale=8.0):
    conf = dual_softmax_conf(feat_c0, feat_c1, temperature)
    out = coarse_match_from_conf(conf, hw0_c, hw1_c, thr, border_rm, scale)
    out["conf_matrix"] = conf
    return out


# --------------------------------------------------------------------------------------- a7
def fine_preprocess(p, feat_f0, feat_f1, feat_c0, feat_c1, b_ids, i_ids, j_ids, W=5, stride=4):
    """mp3d_loftr/src/loftr/loftr_module/fine_preprocess.py:29-59 (fine_concat_coarse_feat=True)."""
    cf = feat_f0.shape[1]
    if b_ids.shape[0] == 0:
        return torch.empty(0, W * W, cf), torch.empty(0, W * W, cf)
    n = feat_f0.shape[0]

    def unfold(f):
        u = F.unfold(f, kernel_size=(W, W), stride=stride, padding=W // 2)  # [n, c*ww, l]
        return u.view(n, cf, W * W, -1).permute(0, 3, 2, 1)  # 'n (c ww) l -> n l ww c'

    f0 = unfold(feat_f0)[b_ids, i_ids]
    f1 = unfold(feat_f1)[b_ids, j_ids]
    c_win = F.linear(torch.cat([feat_c0[b_ids, i_ids], feat_c1[b_ids, j_ids]], 0),
                     p["down_proj.weight"], p["down_proj.bias"])
    cat = torch.cat([torch.cat([f0, f1], 0), c_win[:, None, :].expand(-1, W * W, -1)], -1)
    out = F.linear(cat, p["merge_feat.weight"], p["merge_feat.bias"])
    return torch.chunk(out, 2, dim=0)


# --------------------------------------------------------------------------------------- a8
def fine_matching(feat_f0, feat_f1, mkpts0_c, mkpts1_c, scale=2.0):
    """mp3d_loftr/src/loftr/utils/fine_matching.py:15-76 (eval; kornia 0.7.1 dsnt restated).

    Returns expec_f [M,3], mkpts0_f, mkpts1_f."""
    M, WW, C = feat_f0.shape
    W = int(math.sqrt(WW))
    if M == 0:
        return torch.empty(0, 3), mkpts0_c, mkpts1_c
    picked = feat_f0[:, WW // 2, :]
    sim = torch.einsum("mc,mrc->mr", picked, feat_f1)
    heat = torch.softmax(sim / C ** .5, dim=1)  # [M, WW]
    lin = torch.linspace(-1, 1, W)  # create_meshgrid(W, W, True): x = linspace(-1,1,W) along width
    gx = lin[None, :].expand(W, W).reshape(-1)
    gy = lin[:, None].expand(W, W).reshape(-1)
    grid = torch.stack([gx, gy], -1)  # [WW, 2]
    coords = torch.stack([(heat * gx).sum(-1), (heat * gy).sum(-1)], -1)  # spatial_expectation2d
    var = torch.sum(grid[None] ** 2 * heat[:, :, None], dim=1) - coords ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    expec_f = torch.cat([coords, std[:, None]], -1)
    mkpts1_f = mkpts1_c + (coords * (W // 2) * scale)[:len(mkpts1_c)]
    return expec_f, mkpts0_c, mkpts1_f


# --------------------------------------------------------------------------------------- backbone (inside the boundary)
def _bn(x, p, name):
    return F.batch_norm(x, p[name + ".running_mean"], p[name + ".running_var"], p[name + ".weight"],
                        p[name + ".bias"], False, 0.0, 1e-5)


def _basic_block(p, x, stride):
    """mp3d_loftr/src/loftr/backbone/resnet_fpn.py:15-40."""
    y = F.relu(_bn(F.conv2d(x, p["conv1.weight"], None, stride, 1), p, "bn1"))
    y = _bn(F.conv2d(y, p["conv2.weight"], None, 1, 1), p, "bn2")
    if stride != 1:
        x = _bn(F.conv2d(x, p["downsample.0.weight"], None, stride, 0), p, "downsample.1")
    return F.relu(x + y)


def resnet_fpn_8_2(p, x):
    """resnet_fpn.py:101-119 (eval-mode BatchNorm).  x [N,1,H,W] -> (feat_c [N,256,H/8,W/8], feat_f [N,128,H/2,W/2])."""
    x0 = F.relu(_bn(F.conv2d(x, p["conv1.weight"], None, 2, 3), p, "bn1"))
    x1 = _basic_block(_sub(p, "layer1.1"), _basic_block(_sub(p, "layer1.0"), x0, 1), 1)
    x2 = _basic_block(_sub(p, "layer2.1"), _basic_block(_sub(p, "layer2.0"), x1, 2), 1)
    x3 = _basic_block(_sub(p, "layer3.1"), _basic_block(_sub(p, "layer3.0"), x2, 2), 1)
    x3_out = F.conv2d(x3, p["layer3_outconv.weight"])
    x3_out_2x = F.interpolate(x3_out, scale_factor=2., mode="bilinear", align_corners=True)
    x2_out = F.conv2d(x2, p["layer2_outconv.weight"])

    def outconv2(pp, t):
        t = F.conv2d(t, pp["0.weight"], None, 1, 1)
        t = F.leaky_relu(_bn(t, pp, "1"), 0.01)
        return F.conv2d(t, pp["3.weight"], None, 1, 1)

    x2_out = outconv2(_sub(p, "layer2_outconv2"), x2_out + x3_out_2x)
    x2_out_2x = F.interpolate(x2_out, scale_factor=2., mode="bilinear", align_corners=True)
    x1_out = F.conv2d(x1, p["layer1_outconv.weight"])
    x1_out = outconv2(_sub(p, "layer1_outconv2"), x1_out + x2_out_2x)
    return x3_out, x1_out


# --------------------------------------------------------------------------------------- LoFTR.forward
def loftr_forward(sd, image0, image1, cfg):
    """mp3d_loftr/src/loftr/loftr.py:56-135 (forward_feature_extraction + forward_correspondence_prediction),
    eval, equal image sizes, no masks.  `cfg` is the lower-cased dict (oracle/ref_import.mp3d_eval_config layout).
    Returns the dict of keys the reference writes into `data`."""
    bs = image0.shape[0]
    feats_c, feats_f = resnet_fpn_8_2(_sub(sd, "backbone"), torch.cat([image0, image1], 0))
    feat_c0, feat_c1 = feats_c.split(bs)
    feat_f0, feat_f1 = feats_f.split(bs)
    data = {"bs": bs, "hw0_i": tuple(image0.shape[2:]), "hw1_i": tuple(image1.shape[2:]),
            "hw0_c": tuple(feat_c0.shape[2:]), "hw1_c": tuple(feat_c1.shape[2:]),
            "hw0_f": tuple(feat_f0.shape[2:]), "hw1_f": tuple(feat_f1.shape[2:]),
            "featmap_f0": feat_f0, "featmap_f1": feat_f1, "feats_c": feats_c}
    tb = cfg["coarse"]["temp_bug_fix"]
    fc0 = add_pos_and_flatten(feat_c0, tb)
    fc1 = add_pos_and_flatten(feat_c1, tb)
    fc0, fc1 = local_feature_transformer(_sub(sd, "loftr_coarse"), fc0, fc1, cfg["coarse"]["layer_names"],
                                         cfg["coarse"]["nhead"])
    mc = cfg["match_coarse"]
    scale = data["hw0_i"][0] / data["hw0_c"][0]
    data.update(coarse_matching(fc0, fc1, data["hw0_c"], data["hw1_c"], mc["thr"], mc["border_rm"],
                                mc["dsmax_temperature"], scale))
    W = cfg["fine_window_size"]
    stride = data["hw0_f"][0] // data["hw0_c"][0]
    ff0, ff1 = fine_preprocess(_sub(sd, "fine_preprocess"), feat_f0, feat_f1, fc0, fc1, data["b_ids"],
                               data["i_ids"], data["j_ids"], W, stride)
    if ff0.size(0) != 0:
        ff0, ff1 = local_feature_transformer(_sub(sd, "loftr_fine"), ff0, ff1, cfg["fine"]["layer_names"],
                                             cfg["fine"]["nhead"])
    expec_f, mk0, mk1 = fine_matching(ff0, ff1, data["mkpts0_c"], data["mkpts1_c"],
                                      data["hw0_i"][0] / data["hw0_f"][0])
    data.update({"W": W, "expec_f": expec_f, "mkpts0_f": mk0, "mkpts1_f": mk1, "featmap0": fc0, "featmap1": fc1,
                 "mask_c0": None, "mask_c1": None, "translation_scale": None})
    return data


# --------------------------------------------------------------------------------------- a10
def normalize_points(points, eps=1e-8):
    """*/third_party/prior_ransac/cv_geometry.py:713-750.  points [B,N,2] -> (points_norm, T [B,3,3])."""
    x_mean = points.mean(dim=1, keepdim=True)
    scale = (points - x_mean).norm(dim=-1, p=2).mean(dim=-1)
    scale = torch.sqrt(torch.tensor(2.0)) / (scale + eps)
    ones, zeros = torch.ones_like(scale), torch.zeros_like(scale)
    T = torch.stack([scale, zeros, -scale * x_mean[..., 0, 0], zeros, scale, -scale * x_mean[..., 0, 1],
                     zeros, zeros, ones], dim=-1).view(-1, 3, 3)
    # transform_points (linalg.py): homogeneous multiply + divide by w (w == 1 here)
    ph = F.pad(points, [0, 1], value=1.0)
    pn = ph @ T.transpose(-2, -1)
    pn = pn[..., :2] / pn[..., 2:]
    return pn, T


def _svd(x):
    """torch_utils.py:13-33 `_torch_svd_cast`: returns (U, S, V) with V = Vh^H."""
    U, S, Vh = torch.linalg.svd(x)
    return U, S, Vh.transpose(-2, -1)


def run_8point(points1, points2, weights=None, dense_diag=False):
    """cv_geometry.py:772-833: weighted normalised 8-point, rank-2 projection, de-normalise, /F22.

    `dense_diag=True` reproduces the reference's literal `X^T diag_embed(w) X` (:813-817, O(N^2) memory);
    the default uses the algebraically identical (w[:,None]*X)^T X so large N fits in memory."""
    assert points1.shape == points2.shape and points1.shape[1] >= 8
    p1, T1 = normalize_points(points1)
    p2, T2 = normalize_points(points2)
    x1, y1 = torch.chunk(p1, dim=-1, chunks=2)
    x2, y2 = torch.chunk(p2, dim=-1, chunks=2)
    ones = torch.ones_like(x1)
    X = torch.cat([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, ones], dim=-1)  # :810
    if weights is None:
        A = X.transpose(-2, -1) @ X
    elif dense_diag:
        A = X.transpose(-2, -1) @ torch.diag_embed(weights) @ X
    else:
        A = X.transpose(-2, -1) @ (weights[..., None] * X)
    _, _, V = _svd(A)
    Fm = V[..., -1].view(-1, 3, 3)
    U, S, V = _svd(Fm)
    rank_mask = torch.tensor([1.0, 1.0, 0.0], dtype=Fm.dtype)
    Fp = U @ (torch.diag_embed(S * rank_mask) @ V.transpose(-2, -1))
    Fe = T2.transpose(-2, -1) @ (Fp @ T1)
    nv = Fe[..., -1:, -1:]  # normalize_transformation :753-769
    return torch.where(nv.abs() > 1e-8, Fe / (nv + 1e-8), Fe)


# --------------------------------------------------------------------------------------- a11
def decompose_essential_matrix(E):
    """*/third_party/prior_ransac/essential.py:99-139 -> (R1, R2, t[...,3,1])."""
    U, _, V = _svd(E)
    Vt = V.transpose(-2, -1)
    mask = torch.ones_like(E)
    mask[..., -1:] *= -1.0
    maskt = mask.transpose(-2, -1)
    U = torch.where((torch.det(U) < 0.0)[..., None, None], U * mask, U)
    Vt = torch.where((torch.det(Vt) < 0.0)[..., None, None], Vt * maskt, Vt)
    W = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], dtype=E.dtype)
    R1 = U @ W @ Vt
    R2 = U @ W.transpose(-2, -1) @ Vt
    return R1, R2, U[..., -1:]


def motion_from_essential(E):
    """essential.py:41-64 -> Rs [*,4,3,3], ts [*,4,3,1] in the order (R1,t),(R1,-t),(R2,t),(R2,-t)."""
    R1, R2, t = decompose_essential_matrix(E)
    return torch.stack([R1, R1, R2, R2], dim=-3), torch.stack([t, -t, t, -t], dim=-3)


# --------------------------------------------------------------------------------------- a15 helpers
def rotation_6d_to_matrix(d6):
    """mp3d_loftr/src/losses/loftr_loss.py:10-29 (Gram-Schmidt, rows b1,b2,b1xb2)."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def compute_normalized_6d(pose_mtx, mean=POSE_MEAN_6D, std=POSE_STD_6D):
    """loftr_loss.py:31-39: [t | R[0,:] | R[1,:]] normalised."""
    r6 = pose_mtx[..., :2, :3].reshape(*pose_mtx.shape[:-2], 6)
    tr = pose_mtx[..., :3, 3]
    v = torch.cat([tr, r6], dim=-1)
    return (v - mean.to(v.device)) / std.to(v.device)


def preprocess_helper(loftr_rt, num_corr, num_before, inl_tight, inl_ultra):
    """mp3d_loftr/src/loftr/loftr.py:137-171 for ONE pair (the reference is batch-1 only, §7):
    loftr_rt [3,4] -> (loftr_preds_6d [1,13], inv_loftr_preds_6d [1,13]) with regress_use_num_corres
    and use_many_ransac_thr both on."""
    rt = loftr_rt.float()
    p6 = compute_normalized_6d(rt).unsqueeze(0)
    rt44 = torch.cat([loftr_rt, torch.tensor([[0, 0, 0, 1.]], dtype=loftr_rt.dtype)], dim=0)
    ip6 = compute_normalized_6d(torch.linalg.inv(rt44)[:3, :4]).float().unsqueeze(0)
    extra = torch.tensor([[num_corr / 500.0, num_before / 500.0, inl_tight / 500.0, inl_ultra / 500.0]],
                         dtype=torch.float32)
    return torch.cat([p6, extra], -1), torch.cat([ip6, extra], -1)


# --------------------------------------------------------------------------------------- a12
def emm_positional_encodings_mp3d():
    """mp3d_loftr/src/loftr/loftr_module/transformer.py:183-248 with its hard-coded h,w=60,80 and intrinsics
    [517/9, 517/8, 40, 30] (:194-196) => a constant [4800, 6] table (y^2, x^2, xy, y, x, 1)."""
    h, w = 60, 80
    fx, fy, cx, cy = (torch.tensor(v) for v in (517 / 9, 517 / 8, 40., 30.))
    ys = torch.linspace(-1, 1, steps=h)
    xs = torch.linspace(-1, 1, steps=w)
    hpix, wpix = cy * 2, cx * 2
    K = torch.zeros(3, 3)
    K[0, 0] = (fx / wpix) * 2
    K[1, 1] = (fy / hpix) * 2
    K[0, 2] = (cx / wpix) * 2 - 1
    K[1, 2] = (cy / hpix) * 2 - 1
    K[2, 2] = 1
    Kinv = torch.inverse(K)
    jj, kk = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    pts = torch.stack([xs[kk.reshape(-1)], ys[jj.reshape(-1)], torch.ones(h * w)], 0)  # [3, hw], index j*w+k
    wv = Kinv @ pts
    p4 = wv[0] / wv[2]  # x
    p3 = wv[1] / wv[2]  # y
    return torch.stack([p3 * p3, p4 * p4, p3 * p4, p3, p4, torch.ones(h * w)], dim=1)


def emm_positional_encodings_vit(intrinsics, h=24, w=24):
    """interiornetStreetlearn_8ptVit/src/modules/vision_transformer.py:90-158 for intrinsics [B,2,4]
    (already scaled to the feature grid); index quirk `k*w+j` (:150-151) kept."""
    B = intrinsics.shape[0]
    ys = torch.linspace(-1, 1, steps=h)
    xs = torch.linspace(-1, 1, steps=w)
    fx, fy, cx, cy = intrinsics[:, 0].unbind(dim=-1)
    hpix, wpix = cy * 2, cx * 2
    K = torch.zeros(B, 3, 3)
    K[:, 0, 0] = (fx / wpix) * 2
    K[:, 1, 1] = (fy / hpix) * 2
    K[:, 0, 2] = (cx / wpix) * 2 - 1
    K[:, 1, 2] = (cy / hpix) * 2 - 1
    K[:, 2, 2] = 1
    Kinv = torch.inverse(K)
    p3 = torch.zeros(B, h * w)
    p4 = torch.zeros(B, h * w)
    for j in range(h):
        for k in range(w):
            wv = Kinv @ torch.tensor([xs[k], ys[j], 1.0])
            p3[:, k * w + j] = wv[:, 1] / wv[:, 2]
            p4[:, k * w + j] = wv[:, 0] / wv[:, 2]
    return torch.stack([p3 * p3, p4 * p4, p3 * p4, p3, p4, torch.ones(B, h * w)], dim=2)


# --------------------------------------------------------------------------------------- a13
def cross_attention_emm(p, x1, x2, positional, num_heads):
    """transformer.py:266-303 / vision_transformer.py:177-208: dual-softmax bilinear attention.
    x1,x2 [B,N,C]; positional [B or 1,N,6].  Returns (fundamental_2, fundamental_1) each [B, d+6, C]."""
    B, N, C = x1.shape
    d = C // num_heads
    scale = d ** -0.5

    def qkv(x):
        t = F.linear(x, p["qkv.weight"], p["qkv.bias"]).reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
        return t[0], t[1], t[2]

    q1, k1, v1 = qkv(x1)
    q2, k2, v2 = qkv(x2)
    attn_1 = (q2 @ k1.transpose(-2, -1)) * scale
    attn_2 = (q1 @ k2.transpose(-2, -1)) * scale
    af1 = attn_1.softmax(dim=-1) * attn_1.softmax(dim=-2)
    af2 = attn_2.softmax(dim=-1) * attn_2.softmax(dim=-2)
    pos = positional.expand(B, N, 6).unsqueeze(1).repeat(1, num_heads, 1, 1)
    v1 = torch.cat([v1, pos], dim=3)
    v2 = torch.cat([v2, pos], dim=3)
    f1 = (v1.transpose(-2, -1) @ af1) @ v1
    f2 = (v2.transpose(-2, -1) @ af2) @ v2
    ch = C + 6 * num_heads
    f1 = f1.reshape(B, ch, ch // num_heads).transpose(-2, -1)
    f2 = f2.reshape(B, ch, ch // num_heads).transpose(-2, -1)
    f2 = F.linear(f2, p["proj_fundamental.weight"], p["proj_fundamental.bias"])
    f1 = F.linear(f1, p["proj_fundamental.weight"], p["proj_fundamental.bias"])
    return f2, f1


def _mlp_gelu(p, x):
    """vit_layers/mlp.py:11-27 (exact-erf GELU)."""
    return F.linear(F.gelu(F.linear(x, p["fc1.weight"], p["fc1.bias"])), p["fc2.weight"], p["fc2.bias"])


# --------------------------------------------------------------------------------------- a14
def cross_block_mp3d(p, x, positional, num_heads=4, eps=1e-5):
    """transformer.py:335-348.  x [2,4800,256] = cat(feat0, feat1) of ONE pair (B=1 semantics, §7)."""
    b_s, h_w, nf = x.shape
    if "pos_embed" in p:
        x = x + p["pos_embed"]
    x = x.reshape(-1, 2, h_w, nf)
    n1 = F.layer_norm(x[:, 0], (nf,), p["norm1.weight"], p["norm1.bias"], eps)
    n2 = F.layer_norm(x[:, 1], (nf,), p["norm1.weight"], p["norm1.bias"], eps)
    f1, f2 = cross_attention_emm(_sub(p, "cross_attn"), n1, n2, positional, num_heads)
    fund = torch.cat([f1.unsqueeze(1), f2.unsqueeze(1)], dim=1).reshape(b_s, -1, nf)
    return fund + _mlp_gelu(_sub(p, "mlp"), F.layer_norm(fund, (nf,), p["norm2.weight"], p["norm2.bias"], eps))


def _seq(p, x, idxs, acts):
    for i, a in zip(idxs, acts):
        x = F.linear(x, p[f"{i}.weight"], p[f"{i}.bias"])
        if a == "relu":
            x = F.relu(x)
        elif a == "sigmoid":
            x = torch.sigmoid(x)
    return x


def far_head_mp3d(p, feat0, feat1, loftr_preds, inv_loftr_preds, cfg):
    """LocalFeatureTransformerRegressor.forward + forward_emm, transformer.py:423-499, for ONE pair
    (feat0, feat1 [1,4800,256]; loftr_preds [1,13]), config of record (use_simple_moe, use_2wt, scale_8pt).
    Returns (pose_preds [1,9], pred_RT_wt [1,2])."""
    if cfg["regress_loftr_layers"] > 0:
        feat0, feat1 = local_feature_transformer(_sub(p, "loftr"), feat0, feat1, cfg["regress"]["layer_names"],
                                                 cfg["regress"]["nhead"])
    B = feat0.shape[0]
    x = torch.cat([feat0, feat1], dim=0)
    x = cross_block_mp3d(_sub(p, "emm"), x, emm_positional_encodings_mp3d()[None])
    features = F.layer_norm(x, (x.shape[-1],), p["norm.weight"], p["norm.bias"], 1e-6).reshape(B, -1)
    feats = _seq(_sub(p, "encoder"), features, (0, 2), ("relu", None))
    pred_reg_6d = _seq(_sub(p, "pose_regressor_simple_moe"), feats, (0, 2), ("relu", None))
    pred_reg_t = pred_reg_6d[..., :3]
    loftr_t_in = loftr_preds[..., :3]
    mean, std = POSE_MEAN_6D.to(loftr_t_in.device), POSE_STD_6D.to(loftr_t_in.device)
    if cfg["regress"]["scale_8pt"]:  # :436-446
        lu = loftr_t_in * std[:3] + mean[:3]
        ru = pred_reg_t * std[:3] + mean[:3]
        lu2 = lu[..., :3] * torch.linalg.norm(ru, dim=-1) / torch.clamp(torch.linalg.norm(lu[..., :3], dim=-1),
                                                                         1e-3, 100)
        loftr_t = (lu2 - mean[:3]) / std[:3]
    else:
        loftr_t = loftr_t_in
    pose_size_in = loftr_preds.shape[-1]
    loftr_R = loftr_preds[..., 3:-(pose_size_in - 9)] if pose_size_in > 9 else loftr_preds[..., 3:]
    feats_preds = torch.cat([features, pred_reg_6d, loftr_preds], dim=-1)
    wt = _seq(_sub(p, "moe_predictor"), feats_preds, (0, 2, 4), ("relu", "relu", "sigmoid"))
    pred_T = wt[..., 0] * pred_reg_t + (1 - wt[..., 0]) * loftr_t  # B=1 broadcasting (:466-467)
    pred_R = wt[..., 1] * pred_reg_6d[..., 3:] + (1 - wt[..., 1]) * loftr_R
    return torch.cat([pred_T, pred_R], dim=-1), wt


def prior_rt_from_regressed(regressed_rt):
    """loftr.py:188-192: de-normalise the head output, Gram-Schmidt, -> priorRT [3,4]."""
    mean, std = POSE_MEAN_6D.to(regressed_rt.device), POSE_STD_6D.to(regressed_rt.device)
    R = regressed_rt[:, 3:] * std[3:] + mean[3:]
    t = regressed_rt[0, :3] * std[:3] + mean[:3]
    R = rotation_6d_to_matrix(R)[0]
    return torch.cat([R, t[:, None]], dim=-1)


# --------------------------------------------------------------------------------------- a16 (8pt-ViT)
def vit_attention(p, x, num_heads):
    """interiornetStreetlearn_8ptVit/src/modules/vision_transformer.py:236-262."""
    B, N, C = x.shape
    d = C // num_heads
    qkv = F.linear(x, p["qkv.weight"], p["qkv.bias"]).reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, p["proj.weight"], p["proj.bias"])


def vit_block(p, x, num_heads, eps=1e-6):
    """vision_transformer.py:265-283 (pre-norm)."""
    C = x.shape[-1]
    x = x + vit_attention(_sub(p, "attn"), F.layer_norm(x, (C,), p["norm1.weight"], p["norm1.bias"], eps), num_heads)
    return x + _mlp_gelu(_sub(p, "mlp"), F.layer_norm(x, (C,), p["norm2.weight"], p["norm2.bias"], eps))


def vit_cross_block(p, x, positional, num_heads, eps=1e-6):
    """vision_transformer.py:224-234 (no pos_embed inside the block; x [2B,N,C] interleaved pairs)."""
    b_s, h_w, nf = x.shape
    x = x.reshape(-1, 2, h_w, nf)
    n1 = F.layer_norm(x[:, 0], (nf,), p["norm1.weight"], p["norm1.bias"], eps)
    n2 = F.layer_norm(x[:, 1], (nf,), p["norm1.weight"], p["norm1.bias"], eps)
    f1, f2 = cross_attention_emm(_sub(p, "cross_attn"), n1, n2, positional, num_heads)
    fund = torch.cat([f1.unsqueeze(1), f2.unsqueeze(1)], dim=1).reshape(b_s, -1, nf)
    return fund + _mlp_gelu(_sub(p, "mlp"), F.layer_norm(fund, (nf,), p["norm2.weight"], p["norm2.bias"], eps))


def compute_rotation_matrix_from_ortho6d(poses):
    """interiornetStreetlearn_8ptVit/src/geom/RotationContinuity/sanity_test/code/tools.py:47-62
    (columns x, y, z; differs from rotation_6d_to_matrix's row stacking)."""
    x_raw, y_raw = poses[:, 0:3], poses[:, 3:6]

    def nrm(v):  # normalize_vector :20-28 : v / max(|v|, 1e-8)
        mag = torch.sqrt((v ** 2).sum(1))
        mag = torch.max(mag, torch.tensor(1e-8))
        return v / mag[:, None]

    x = nrm(x_raw)
    z = nrm(torch.cross(x, y_raw, dim=1))
    y = torch.cross(z, x, dim=1)
    return torch.cat((x[:, :, None], y[:, :, None], z[:, :, None]), 2)


def vit_fusion_head(sd, feats, positional, loftr_preds, loftr_num_corr, pose_mean, pose_std, depth=6, num_heads=3):
    """interiornetStreetlearn_8ptVit/src/model.py:170-217 after extract_features.
    feats [2B,576,192] (post extractor, token-major).  Returns (t [B,3], R [B,3,3], r6d [B,6], w [B,2])."""
    B = feats.shape[0] // 2
    ft = _sub(sd, "fusion_transformer")
    x = feats + ft["pos_embed"]
    for layer in range(depth - 1):
        x = vit_block(_sub(ft, f"blocks.{layer}"), x, num_heads)
    x = vit_cross_block(_sub(ft, f"blocks.{depth - 1}"), x, positional, num_heads)
    features = F.layer_norm(x, (x.shape[-1],), ft["norm.weight"], ft["norm.bias"], 1e-6).reshape(B, -1)
    lp = loftr_preds.float()
    r6 = lp[..., :2, :3].reshape(B, 6)
    l6 = (torch.cat([lp[..., :3, 3], r6], -1) - pose_mean) / pose_std
    l6 = torch.cat([l6, loftr_num_corr.float().unsqueeze(1) / 500], dim=-1)
    pred = _seq(_sub(sd, "pose_regressor"), features, (0, 2, 4), ("relu", "relu", None))
    wt = _seq(_sub(sd, "moe_predictor"), torch.cat([features, pred, l6], -1), (0, 2, 4), ("relu", "relu", "sigmoid"))
    pT = wt[..., :1] * pred[..., :3] + (1 - wt[..., :1]) * l6[..., :3]
    pR = wt[..., 1:] * pred[..., 3:] + (1 - wt[..., 1:]) * l6[..., 3:-1]
    r6d = pR
    R = compute_rotation_matrix_from_ortho6d(r6d * pose_std[3:] + pose_mean[3:])
    t = pT * pose_std[:3] + pose_mean[:3]
    return t, R, r6d, wt


# --------------------------------------------------------------------------------------- a17 (mapfree)
def mapfree_regression_mlp(sd, feats, loftr_rt, inliers3):
    """mapfree_6dreg/lib/models/regression/model.py:198-233.  feats [B,27648]; loftr_rt [B,3,4];
    inliers3 [B,3] raw counts.  Returns (R6d [B,6], t [B,3], w [B,2])."""
    B = feats.shape[0]
    lp = loftr_rt.float()
    l9 = torch.cat([lp[..., :3, 3], lp[..., :2, :3].reshape(B, 6)], -1)  # compute_6d, lib/utils/loss.py:9
    pred = _seq(_sub(sd, "pose_regressor"), feats, (0, 2, 4), ("relu", "relu", None))
    ratio = torch.linalg.norm(pred[..., :3], dim=-1) / torch.clamp(torch.linalg.norm(l9[..., :3], dim=-1), 1e-2, 1e2)
    lt = l9[..., :3] * torch.clamp(ratio.unsqueeze(1), 1e-2, 1e2)  # :221-223
    inl = inliers3.float() / 500  # use_prior branch (:207)
    lout = torch.cat([lt, l9[..., 3:], inl], -1)
    wt = _seq(_sub(sd, "moe_predictor"), torch.cat([feats, pred, lout], -1), (0, 2, 4),
              ("relu", "relu", "sigmoid"))
    t = wt[..., :1] * pred[..., :3] + (1 - wt[..., :1]) * lout[..., :3]
    R = wt[..., 1:] * pred[..., 3:] + (1 - wt[..., 1:]) * lout[..., 3:-3]
    return R, t, wt


# --------------------------------------------------------------------------------------- solver glue used by bench/tests
def pose_from_matches_8pt(mk0, mk1, w, K0, K1):
    """Config-2/5 solver: K-normalise keypoints (metrics.py:88-89), weighted run_8point
    (weights = mconf), then the 4 (R,t) candidates of decompose_essential_matrix.  On K-normalised points the
    8-point result IS the essential matrix.  <8 points -> identity pose (metrics.py:83-85 analogue).
    Candidate choice = cheirality vote (positive depth in both cameras by linear triangulation), the same
    criterion cv2.recoverPose applies (metrics.py:165); ties -> lowest candidate index."""
    if mk0.shape[0] < 8:
        return torch.eye(3), torch.zeros(3), torch.eye(3)
    k0 = (mk0 - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    k1 = (mk1 - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    E = run_8point(k0[None].float(), k1[None].float(), w[None].float())[0]
    Rs, ts = motion_from_essential(E)
    best, bi = -1, 0
    for c in range(4):
        n = cheirality_count(Rs[c], ts[c, :, 0], k0.float(), k1.float())
        if n > best:
            best, bi = n, c
    return Rs[bi], ts[bi, :, 0], E


def cheirality_count(R, t, x0, x1):
    """#points with positive depth in both views for x1 ~ R x0 + t (normalised coords); depth via the
    standard two-view mid-point-free linear solve  z0 * (R x0h) x x1h = -t x x1h."""
    x0h = F.pad(x0, [0, 1], value=1.0)
    x1h = F.pad(x1, [0, 1], value=1.0)
    a = torch.cross(x0h @ R.T, x1h, dim=-1)
    b = torch.cross(t[None].expand_as(x1h), x1h, dim=-1)
    z0 = -(a * b).sum(-1) / (a * a).sum(-1).clamp_min(1e-20)
    X1 = z0[:, None] * (x0h @ R.T) + t[None]
    return int(((z0 > 0) & (X1[:, 2] > 0)).sum())


# --------------------------------------------------------------------------------------- deterministic synthetic inputs
# --------------------------------------------------------------------------------------- 8f rank 2: prior-guided RANSAC
def sampson_epipolar_distance(pts1, pts2, Fm):
    """kornia 0.7.1 `sampson_epipolar_distance(..., squared=True)` (un-vendored; restated from its public definition,
    used at ransac.py:151,256-268):  (p2^T F p1)^2 / (|(F p1)_{1:2}|^2 + |(F^T p2)_{1:2}|^2).  pts [*,N,2], Fm [*,3,3]."""
    p1 = F.pad(pts1, [0, 1], value=1.0)
    p2 = F.pad(pts2, [0, 1], value=1.0)
    l1in2 = p1 @ Fm.transpose(-2, -1)
    l2in1 = p2 @ Fm
    num = (p2 * l1in2).sum(-1).pow(2)
    den = l1in2[..., :2].norm(2, dim=-1).pow(2) + l2in1[..., :2].norm(2, dim=-1).pow(2)
    return num / den


def symmetrical_epipolar_distance(pts1, pts2, Fm):
    """kornia 0.7.1 `symmetrical_epipolar_distance(..., squared=True)` (ransac.py:364)."""
    p1 = F.pad(pts1, [0, 1], value=1.0)
    p2 = F.pad(pts2, [0, 1], value=1.0)
    l1in2 = p1 @ Fm.transpose(-2, -1)
    l2in1 = p2 @ Fm
    num = (p2 * l1in2).sum(-1).pow(2)
    return num * (1.0 / l1in2[..., :2].norm(2, dim=-1).pow(2) + 1.0 / l2in1[..., :2].norm(2, dim=-1).pow(2))


def essential_from_prior_rt(RT):
    """ransac.py:63-71 `fundamental_from_RT` (which returns E, not F): kornia `essential_from_Rt(I, 0, R, t)` = [t]_x R."""
    R, t = RT[..., :3, :3], RT[..., :3, 3]
    z = torch.zeros_like(t[..., 0])
    tx = torch.stack([torch.stack([z, -t[..., 2], t[..., 1]], -1), torch.stack([t[..., 2], z, -t[..., 0]], -1),
                      torch.stack([-t[..., 1], t[..., 0], z], -1)], -2)
    return tx @ R


def ransac_bias_weight(kp1, kp2, prior_rt, sigma_sq=0.1):
    """ransac.py:358-367 with use_linear_bias_sampling: exp(-sym_epipolar(kp1, kp2, E_prior) / sigma^2).  prior_rt [3,4]
    with unit-norm translation (setup_prior, ransac.py:180)."""
    return torch.exp(-symmetrical_epipolar_distance(kp1[None], kp2[None], essential_from_prior_rt(prior_rt)[None]) / sigma_sq)[0]


def ransac_prior_estimate(models, prior_rt, pcl, prior_lambda=0.3, both_signs=False):
    """ransac.py:203-231 + :401-404 (use_noexp_prior_scoring): each model E -> (R1, R2, T) by
    decompose_essential_matrix; error_k = mean |[R_k | T] pcl - prior_rt pcl| over the 3 x npcl coordinates;
    prior score = -(min(error_1, error_2))^2 / lambda.  models [H,3,3], prior_rt [3,4], pcl [npcl,3].
    both_signs=True additionally scores -T (the sign of T = U[:,2] is whatever the SVD routine returns; the CUDA
    kernel scores both and keeps the better)."""
    R1, R2, T = decompose_essential_matrix(models)
    target = prior_rt[:, :3] @ pcl.t() + prior_rt[:, 3:]                         # [3, npcl]
    def err(R, Tt):
        return ((R @ pcl.t()[None] + Tt) - target[None]).abs().reshape(models.shape[0], -1).mean(1)
    e = torch.minimum(err(R1, T), err(R2, T))
    if both_signs:
        e = torch.minimum(e, torch.minimum(err(R1, -T), err(R2, -T)))
    return -e ** 2 / prior_lambda


def ransac_good_models(models):
    """ransac.py:303-308 `remove_bad_models`: keep models whose main diagonal has min |.| > 1e-4."""
    return torch.diagonal(models, dim1=1, dim2=2).abs().min(dim=1)[0] > 1e-4


def ransac_verify(kp1, kp2, models, inl_th, prior_score):
    """ransac.py:256-292: Sampson errors of every model on every correspondence, inlier count at inl_th plus the prior
    score, argmax; inlier masks of the best model at inl_th, inl_th/10, inl_th/100.
    kp [N,2], models [H,3,3], prior_score [H] -> (best index, score [H], masks [3,N])."""
    H = models.shape[0]
    errors = sampson_epipolar_distance(kp1[None].expand(H, -1, 2), kp2[None].expand(H, -1, 2), models)
    inl = errors <= inl_th
    score = inl.to(kp1).sum(dim=1) + prior_score.to(kp1)
    best = int(score.argmax())
    masks = torch.stack([inl[best], errors[best] <= inl_th / 10.0, errors[best] <= inl_th / 100.0])
    return best, score, masks


def philox4x32_10(key, c0, c1=0, c2=0, c3=0):
    """Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): counter block -> 4 uint32.
    Vectorised over c0 (numpy uint64 arithmetic).  `key` = 64-bit seed (low word key0, high word key1)."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    c = [np.asarray(x, dtype=np.uint64) & mask for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, -1).astype(np.uint64)


def ransac_sample_indices(weights, H, seed, pair=0, S=8):
    """The sampling contract of far_ransac_sample_models (include/far_sm100.h): for hypothesis h of pair `pair`,
    draw k = searchsorted(cumsum(w), u * sum(w), 'right') with u = philox(seed; (pair*H+h, block, 0, 0))[j] * 2^-32,
    consuming the 4 words of a block in order and a new block when they run out; a draw that repeats an index already
    in the sample is redrawn, at most 4 times.  The distribution is the reference's np.random.choice(p = w / sum w)
    (ransac.py:161-175); the generator is ours.  weights: fp64 numpy [n] (already including the +1e-4).  -> [H,S]."""
    cdf = np.cumsum(np.asarray(weights, dtype=np.float64))
    n = cdf.shape[0]
    out = np.full((H, S), -1, dtype=np.int64)
    if n < S:
        return out
    for h in range(H):
        g = pair * H + h
        blk, used, words = 0, 4, None
        for k in range(S):
            pick = 0
            for _ in range(5):
                if used == 4:
                    words = philox4x32_10(seed, g, blk)
                    blk, used = blk + 1, 0
                target = (float(words[used]) / 4294967296.0) * cdf[-1]
                used += 1
                pick = min(int(np.searchsorted(cdf, target, side='right')), n - 1)
                if pick not in out[h, :k]:
                    break
            out[h, k] = pick
    return out


# --------------------------------------------------------------------------------------- 8f rank 2: 5-point minimal solver
def _mul_deg_one(a, b):
    """kornia 0.7.1 `solvers.multiply_deg_one_poly` (un-vendored; restated): product of two linear forms in (x, y, z, 1)
    -> the 10 coefficients of [x^2, xy, xz, x, y^2, yz, y, z^2, z, 1].  a, b: [..., 4]."""
    return np.stack([a[..., 0] * b[..., 0], a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0],
                     a[..., 0] * b[..., 2] + a[..., 2] * b[..., 0], a[..., 0] * b[..., 3] + a[..., 3] * b[..., 0],
                     a[..., 1] * b[..., 1], a[..., 1] * b[..., 2] + a[..., 2] * b[..., 1],
                     a[..., 1] * b[..., 3] + a[..., 3] * b[..., 1], a[..., 2] * b[..., 2],
                     a[..., 2] * b[..., 3] + a[..., 3] * b[..., 2], a[..., 3] * b[..., 3]], -1)


def _mul_deg_two_one(a, b):
    """kornia 0.7.1 `solvers.multiply_deg_two_one_poly` (restated): degree-2 poly (10 coefficients, order above) times a
    linear form -> 20 coefficients of [x^3, y^3, x^2y, xy^2, x^2z, x^2, y^2z, y^2, xyz, xy | xz^2, xz, x, yz^2, yz, y, z^3,
    z^2, z, 1] -- Nister's ordering: the first ten monomials are eliminated, the last ten are (x, y, 1) x powers of z."""
    return np.stack([
        a[..., 0] * b[..., 0], a[..., 4] * b[..., 1], a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0],
        a[..., 1] * b[..., 1] + a[..., 4] * b[..., 0], a[..., 0] * b[..., 2] + a[..., 2] * b[..., 0],
        a[..., 0] * b[..., 3] + a[..., 3] * b[..., 0], a[..., 4] * b[..., 2] + a[..., 5] * b[..., 1],
        a[..., 4] * b[..., 3] + a[..., 6] * b[..., 1],
        a[..., 1] * b[..., 2] + a[..., 2] * b[..., 1] + a[..., 5] * b[..., 0],
        a[..., 1] * b[..., 3] + a[..., 3] * b[..., 1] + a[..., 6] * b[..., 0],
        a[..., 2] * b[..., 2] + a[..., 7] * b[..., 0],
        a[..., 2] * b[..., 3] + a[..., 3] * b[..., 2] + a[..., 8] * b[..., 0],
        a[..., 3] * b[..., 3] + a[..., 9] * b[..., 0], a[..., 5] * b[..., 2] + a[..., 7] * b[..., 1],
        a[..., 5] * b[..., 3] + a[..., 6] * b[..., 2] + a[..., 8] * b[..., 1],
        a[..., 6] * b[..., 3] + a[..., 9] * b[..., 1], a[..., 7] * b[..., 2],
        a[..., 7] * b[..., 3] + a[..., 8] * b[..., 2], a[..., 8] * b[..., 3] + a[..., 9] * b[..., 2],
        a[..., 9] * b[..., 3]], -1)


def five_point_constraints(null4):
    """cv_geometry.py:901-945 of `run_5point_our_kornia`: E(x,y,z) = x N0 + y N1 + z N2 + N3 with the four null vectors
    null4 [9, 4] (entry (i,j) of the 3x3 is vector index 3j+i, `fun(i,j)` :898-899) -> the [10, 20] coefficient matrix of
    the nine cubic trace constraints  E E^T E - 1/2 tr(E E^T) E = 0  (rows 0-8) and det E = 0 (row 9)."""
    def fun(i, j):
        return null4[3 * j + i]
    c = np.zeros((10, 20))
    c[9] = (_mul_deg_two_one(_mul_deg_one(fun(0, 1), fun(1, 2)) - _mul_deg_one(fun(0, 2), fun(1, 1)), fun(2, 0))
            + _mul_deg_two_one(_mul_deg_one(fun(0, 2), fun(1, 0)) - _mul_deg_one(fun(0, 0), fun(1, 2)), fun(2, 1))
            + _mul_deg_two_one(_mul_deg_one(fun(0, 0), fun(1, 1)) - _mul_deg_one(fun(0, 1), fun(1, 0)), fun(2, 2)))
    d = {}
    for i in range(3):
        for j in range(3):
            d[(i, j)] = sum(_mul_deg_one(fun(i, k), fun(j, k)) for k in range(3))
    tr = 0.5 * (d[(0, 0)] + d[(1, 1)] + d[(2, 2)])
    for i in range(3):
        d[(i, i)] = d[(i, i)] - tr
    cnt = 0
    for i in range(3):
        for j in range(3):
            c[cnt] = sum(_mul_deg_two_one(d[(i, k)], fun(k, j)) for k in range(3))
            cnt += 1
    return c


def run_5point_nister(points1, points2):
    """cv_geometry.py:861-1041 `run_5point_our_kornia` for ONE minimal sample (5 correspondences, unit weights), fp64:
    null space of the 5 epipolar equations (the 4 smallest right-singular vectors of X^T X, :886-896), the 10 x 20
    constraint matrix, Gauss-Jordan on the cubic monomials (:952-956), Nister's 3 x 3 polynomial matrix A(z) (:958-969),
    det A(z) = a degree-10 polynomial in z (the reference expands it with `determinant_to_polynomial`, :23-551; here by
    polynomial products: the same polynomial), its REAL roots (the reference keeps the real part of every companion-matrix
    eigenvalue, :984: complex roots yield junk models that its scoring step discards), x and y from two rows of A (:1008)
    and E = -x N0 - y N1 + z N2 + N3 normalised (:1018-1022), returned as p2^T E p1 = 0 matrices [n_real, 3, 3]."""
    x1, y1 = points1[:, 0].astype(np.float64), points1[:, 1].astype(np.float64)
    x2, y2 = points2[:, 0].astype(np.float64), points2[:, 1].astype(np.float64)
    X = np.stack([x1 * x2, x1 * y2, x1, y1 * x2, y1 * y2, y1, x2, y2, np.ones_like(x1)], -1)
    _, _, Vt = np.linalg.svd(X.T @ X)
    null4 = Vt[-4:].T                                   # [9, 4] = V[:, -4:]
    c = five_point_constraints(null4)
    try:
        B = np.linalg.solve(c[:, :10], c[:, 10:])
    except np.linalg.LinAlgError:
        return np.zeros((0, 3, 3))
    A = np.zeros((3, 13))
    for i in range(3):
        A[i, 1:4] = B[4 + 2 * i, 0:3]
        A[i, 0:3] -= B[5 + 2 * i, 0:3]
        A[i, 5:8] = B[4 + 2 * i, 3:6]
        A[i, 4:7] -= B[5 + 2 * i, 3:6]
        A[i, 9:13] = B[4 + 2 * i, 6:10]
        A[i, 8:12] -= B[5 + 2 * i, 6:10]
    P = [[np.poly1d(A[i, 0:4]), np.poly1d(A[i, 4:8]), np.poly1d(A[i, 8:13])] for i in range(3)]
    det = (P[0][0] * (P[1][1] * P[2][2] - P[1][2] * P[2][1]) - P[0][1] * (P[1][0] * P[2][2] - P[1][2] * P[2][0])
           + P[0][2] * (P[1][0] * P[2][1] - P[1][1] * P[2][0]))
    roots = np.roots(det.coeffs)
    roots = np.sort(roots[np.abs(roots.imag) < 1e-9 * (1 + np.abs(roots.real))].real)
    out = []
    for z in roots:
        Bs = np.array([[np.polyval(A[i, 0:4], z), np.polyval(A[i, 4:8], z)] for i in range(3)])
        bs = np.array([np.polyval(A[i, 8:13], z) for i in range(3)])
        xy = np.linalg.lstsq(Bs, bs, rcond=None)[0]
        e = null4[:, 0] * (-xy[0]) + null4[:, 1] * (-xy[1]) + null4[:, 2] * z + null4[:, 3]
        e = e / np.sqrt(xy[0] ** 2 + xy[1] ** 2 + z * z + 1.0)
        out.append(e.reshape(3, 3).T)
    return np.stack(out) if out else np.zeros((0, 3, 3))


# --------------------------------------------------------------------------------------- 8f rank 3: map-free aggregator
def mapfree_correlation_aggregator(vol0, vol1):
    """mapfree_6dreg/lib/models/regression/aggregator.py:42-116 `CorrelationVolumeWarping.forward` with the shipped
    recipe's flags (config/regression/mapfree/rot6d_trans_with_loftr.yaml:8-11: POSITION_ENCODER, MAX_SCORE_CHANNEL; no
    dustbin / normalisation / CV layers):
        C = softmax_j(vol0^T vol1)                      [B, HW, HW]  (6120^2 at the 360x270 map-free resolution)
        vol1w = vol1 C^T ;  pos = grid C^T (grid = meshgrid(linspace(-1,1,H), linspace(-1,1,W)), 'ij')
        out = cat[vol0, vol1w, pos, max_j C]            [B, 2D+3, H, W]
    A flash-attention-shaped op (scores never need to be materialised): the oracle for next round's kernel."""
    B, D, H, W = vol0.shape
    v0, v1 = vol0.reshape(B, D, H * W), vol1.reshape(B, D, H * W)
    c = torch.softmax(torch.bmm(v0.transpose(1, 2), v1), dim=2)
    v1w = torch.bmm(v1, c.transpose(1, 2))
    u = torch.linspace(-1, 1, H, dtype=vol0.dtype)
    v = torch.linspace(-1, 1, W, dtype=vol0.dtype)
    uu, vv = torch.meshgrid(u, v, indexing="ij")
    grid = torch.stack([uu, vv], dim=0).reshape(2, H * W)[None].repeat(B, 1, 1)
    pos = torch.bmm(grid, c.transpose(1, 2))
    mx = c.max(dim=2, keepdim=True)[0].transpose(1, 2)
    return torch.cat([v0, v1w, pos, mx], dim=1).reshape(B, -1, H, W)


def torch_transformer_encoder(sd, x, num_layers=6, nhead=8, eps=1e-5):
    """mapfree_6dreg/lib/models/regression/model.py:57-61,288-291: `nn.TransformerEncoder(nn.TransformerEncoderLayer(
    d_model=256, nhead=8), num_layers=6)` in eval mode on [S, B, E] tokens (S = 12*9 = 108).  torch defaults: post-norm,
    ReLU, dim_feedforward 2048, softmax(QK^T / sqrt(E/nhead)) self-attention with in_proj / out_proj biases.
    `sd`: state dict of the nn.TransformerEncoder (keys layers.{i}.self_attn.in_proj_weight ...)."""
    S, B, E = x.shape
    hd = E // nhead
    for i in range(num_layers):
        p = _sub(sd, f"layers.{i}")
        qkv = F.linear(x, p["self_attn.in_proj_weight"], p["self_attn.in_proj_bias"])
        q, k, v = (t.reshape(S, B * nhead, hd).transpose(0, 1) for t in qkv.chunk(3, dim=-1))
        att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hd), dim=-1) @ v           # [B*nhead, S, hd]
        att = att.transpose(0, 1).reshape(S, B, E)
        att = F.linear(att, p["self_attn.out_proj.weight"], p["self_attn.out_proj.bias"])
        x = F.layer_norm(x + att, (E,), p["norm1.weight"], p["norm1.bias"], eps)
        ff = F.linear(F.relu(F.linear(x, p["linear1.weight"], p["linear1.bias"])), p["linear2.weight"], p["linear2.bias"])
        x = F.layer_norm(x + ff, (E,), p["norm2.weight"], p["norm2.bias"], eps)
    return x


def rng(seed):
    return np.random.default_rng(seed)


def randn(g, *shape, scale=1.0):
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))


def rand(g, *shape):
    return torch.from_numpy(g.random(shape, dtype=np.float32))
