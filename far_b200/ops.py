"""Tensor-level wrappers over the C ABI (include/far_sm100.h).  Each function cites the reference op it replaces.

All inputs must be CUDA tensors; outputs are freshly allocated CUDA tensors.  No torch compute happens here
beyond allocation and (cheap) layout normalisation.
"""
import math
from ctypes import byref

import torch

from . import _lib as L
from ._lib import ptr, f32c, stream, check


def _ws(nbytes, device):
    return L.workspace.get(int(nbytes), device)


class OpTimer:
    """Optional per-C-ABI-call device timing (CUDA events on the launching stream), used by bench.py for the
    `roofline` entry.  Disabled (None) by default: zero overhead on the product path."""

    def __init__(self):
        self.events = {}

    def record(self, name):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.events.setdefault(name, []).append((s, e))
        return s, e

    def summary(self):
        """name -> (calls, total_ms); call after torch.cuda.synchronize()."""
        return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in self.events.items()}


timer = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if timer is not None:
            self.s, self.e = timer.record(self.name)
            self.s.record()

    def __exit__(self, *a):
        if timer is not None:
            self.e.record()
        return False


def linear(x, weight, bias=None, act=L.ACT_NONE, act_cols=-1, x2=None, engine=L.ENGINE_AUTO, out=None):
    """y = act([x | x2] @ weight.T + bias)  -- nn.Linear (transformer.py:23-38 etc.).  x: [..., K1], x2: [..., K2]."""
    lib = L.load()
    lead = x.shape[:-1]
    K1 = x.shape[-1]
    x_ = f32c(x).reshape(-1, K1)
    M = x_.shape[0]
    N, Kt = weight.shape
    K2 = 0
    x2_ = None
    if x2 is not None:
        K2 = x2.shape[-1]
        x2_ = f32c(x2).reshape(-1, K2)
    assert Kt == K1 + K2, (Kt, K1, K2)
    w = f32c(weight)
    b = f32c(bias) if bias is not None else None
    y = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=x.device)
    nws = lib.far_linear_workspace_bytes(M, N, Kt)
    ws = _ws(nws, x.device)
    with _timed("far_linear"):
        check(lib.far_linear(ptr(x_), K1, K1, ptr(x2_), K2 if x2_ is not None else 0, K2, ptr(w), Kt, ptr(b), ptr(y), N,
                           M, N, act, act_cols, engine, ptr(ws), ws.numel(), stream()), "far_linear")
    return y.reshape(*lead, N)


def linear_cat_tail(x, tail, lin, act):
    """act([x | tail] @ lin.weight.T + lin.bias) for a wide feature block `x` [B,K1] and a short, unaligned tail [B,K2]
    (the FAR gates: weight [512, 35840 + 22]).  The weight row pitch K1+K2 is not a multiple of 4 floats, which would
    force the scalar load path over 73 MB of weights; a zero-padded copy (pitch rounded up to 4) is cached ON the
    nn.Linear module, keyed by the weight's version counter and storage pointer."""
    weight, bias = lin.weight, lin.bias
    K1, K2 = x.shape[-1], tail.shape[-1]
    Kp = (K1 + K2 + 3) // 4 * 4
    key = (weight.data_ptr(), weight._version, Kp, str(weight.device))
    cached = getattr(lin, "_far_padded", None)
    if cached is None or cached[0] != key:
        wp = torch.zeros((weight.shape[0], Kp), dtype=torch.float32, device=weight.device)
        wp[:, :K1 + K2] = weight.detach()
        cached = (key, wp)
        lin._far_padded = cached
    tp = torch.zeros((tail.shape[0], Kp - K1), dtype=torch.float32, device=tail.device)
    tp[:, :K2] = tail
    return linear(x, cached[1], bias, act, x2=tp)


def layernorm(x, gamma, beta, eps, residual=None, pre_add=None):
    """nn.LayerNorm over the last dim (transformer.py:59,64-66): residual + LN(x + pre_add).  `pre_add` is a
    [rows_p, C] table broadcast over x's rows modulo rows_p (CrossBlock.pos_embed, transformer.py:337)."""
    lib = L.load()
    C = x.shape[-1]
    x_ = f32c(x).reshape(-1, C)
    r_ = f32c(residual).reshape(-1, C) if residual is not None else None
    p_ = f32c(pre_add).reshape(-1, C) if pre_add is not None else None
    y = torch.empty_like(x_)
    check(lib.far_layernorm_pre(ptr(x_), ptr(p_), p_.shape[0] if p_ is not None else 1, ptr(f32c(gamma)),
                                ptr(f32c(beta)), ptr(r_), ptr(y), x_.shape[0], C, float(eps), stream()),
          "far_layernorm_pre")
    return y.reshape(x.shape)


def pos_encode_flatten(feat, pe_hwc):
    """PositionEncodingSine.forward + rearrange 'n c h w -> n (h w) c' (position_encoding.py:37-42, loftr.py:100-101).
    feat may be NCHW-contiguous or channels_last; pe_hwc: [H*W, C]."""
    lib = L.load()
    if feat.dtype != torch.float32:
        feat = feat.float()
    n, c, h, w = feat.shape
    sn, sc, sh, sw = feat.stride()
    if not (sc == 1 or sw == 1):
        feat = feat.contiguous()
        sn, sc, sh, sw = feat.stride()
    out = torch.empty((n, h * w, c), dtype=torch.float32, device=feat.device)
    check(lib.far_pos_encode_flatten(ptr(feat), sn, sc, sh, sw, ptr(f32c(pe_hwc)), ptr(out), n, c, h, w, stream()),
          "far_pos_encode_flatten")
    return out


def upsample2x_add(low, skip=None):
    """skip + F.interpolate(low, scale_factor=2., mode='bilinear', align_corners=True) on channels_last maps
    (resnet_fpn.py:106-112).  Returns a channels_last [N,C,2H,2W] tensor."""
    lib = L.load()
    n, c, h, w = low.shape
    low_ = f32c(low.permute(0, 2, 3, 1))          # no copy when low is channels_last
    skip_ = f32c(skip.permute(0, 2, 3, 1)) if skip is not None else None
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.float32, device=low.device)
    with _timed("far_upsample2x_add_nhwc"):
        check(lib.far_upsample2x_add_nhwc(ptr(low_), ptr(skip_), ptr(out), n, h, w, c, stream()),
              "far_upsample2x_add_nhwc")
    return out.permute(0, 3, 1, 2)


def stem_conv7x7s2_relu(x, weight, bias):
    """relu(conv2d(x, weight, stride=2, padding=3) + bias) for the 1-channel 7x7 stem (resnet_fpn.py:52-54,80, BN
    folded).  x [N,1,H,W] -> channels_last [N,128,OH,OW]."""
    lib = L.load()
    n, c, h, w = x.shape
    if c != 1 or tuple(weight.shape[1:]) != (1, 7, 7):
        raise L.FarError("stem_conv7x7s2_relu: needs a 1-channel input and a [Cout,1,7,7] kernel")
    cout = weight.shape[0]
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty((n, oh, ow, cout), dtype=torch.float32, device=x.device)
    ws = _ws(lib.far_stem_conv_workspace_bytes(cout), x.device)
    with _timed("far_stem_conv7x7s2_relu_nhwc"):
        check(lib.far_stem_conv7x7s2_relu_nhwc(ptr(f32c(x)), ptr(f32c(weight).reshape(cout, 49)), ptr(f32c(bias)),
                                               ptr(out), n, h, w, cout, ptr(ws), ws.numel(), stream()),
              "far_stem_conv7x7s2_relu_nhwc")
    return out.permute(0, 3, 1, 2)


def scale_shift_act_(x, scale, shift, negative_slope):
    """In place leaky_relu(x*scale[c] + shift[c]) on a channels_last [N,C,H,W] map (eval BatchNorm2d + LeakyReLU,
    resnet_fpn.py:84-95)."""
    lib = L.load()
    n, c, h, w = x.shape
    x_ = x.permute(0, 2, 3, 1)
    if not x_.is_contiguous():
        raise L.FarError("scale_shift_act_ needs a dense channels_last tensor")
    with _timed("far_scale_shift_act_nhwc"):
        check(lib.far_scale_shift_act_nhwc(ptr(x_), ptr(f32c(scale)) if scale is not None else None, ptr(f32c(shift)),
                                           n * h * w, c, float(negative_slope), stream()), "far_scale_shift_act_nhwc")
    return x


def vit_preprocess(images, out_size=224, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """ViTEss.extract_features input preprocessing (interiornetStreetlearn_8ptVit/src/model.py:131-141) in one kernel:
    images [n,3,H,W] fp32 BGR 0..255 (contiguous) -> [n,3,out,out] normalised RGB, nearest resize; bit-identical to
    `F.interpolate((images[:, [2,1,0]] / 255 - mean) / std, size=out)`."""
    import ctypes
    lib = L.load()
    if not (images.is_cuda and images.dtype == torch.float32 and images.is_contiguous() and images.dim() == 4
            and images.shape[1] == 3):
        raise L.FarError("vit_preprocess needs a contiguous fp32 CUDA [n,3,H,W] tensor")
    n, _, H, W = images.shape
    out = torch.empty((n, 3, out_size, out_size), dtype=torch.float32, device=images.device)
    m3, s3 = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    check(lib.far_vit_preprocess(ptr(images), ptr(out), n, H, W, out_size, out_size, m3, s3, stream()),
          "far_vit_preprocess")
    return out


def scale_shift_act(x, scale, shift, negative_slope):
    """Out-of-place leaky_relu(x*scale[c] + shift[c]) on a channels_last [N,C,H,W] map: the pre-activation
    relu(bn(x)) of the map-free PreAct blocks (encoder/preact.py:35,68), x stays alive as the identity shortcut."""
    lib = L.load()
    n, c, h, w = x.shape
    x_ = x.permute(0, 2, 3, 1)
    if not x_.is_contiguous() or x.dtype != torch.float32:
        raise L.FarError("scale_shift_act needs a dense fp32 channels_last tensor")
    y = torch.empty_like(x)
    with _timed("far_scale_shift_act_nhwc"):
        check(lib.far_scale_shift_act_nhwc_out(ptr(x_), ptr(y), ptr(f32c(scale)) if scale is not None else None,
                                               ptr(f32c(shift)), n * h * w, c, float(negative_slope), stream()),
              "far_scale_shift_act_nhwc_out")
    return y


def linear_attention(q, k, v, eps=1e-6, feature_map_applied=False):
    """LinearAttention.forward (linear_attention.py:20-52).  q [N,L,H,D]; k,v [N,S,H,D] -> [N,L,H,D]."""
    lib = L.load()
    N, Lq, H, D = q.shape
    S = k.shape[1]
    q_, k_, v_ = f32c(q), f32c(k), f32c(v)
    out = torch.empty_like(q_)
    nws = lib.far_linear_attention_workspace_bytes(N, S, H, D)
    ws = _ws(nws, q.device)
    C = H * D
    with _timed("far_linear_attention"):
        check(lib.far_linear_attention(ptr(q_), C, ptr(k_), C, ptr(v_), C, ptr(out), C, N, Lq, S, H, D, float(eps),
                                     int(feature_map_applied), ptr(ws), ws.numel(), stream()), "far_linear_attention")
    return out


def tc_weight_split(weight):
    """tcgen05 operand form (tf32 hi | lo) of a static [N,K] weight, computed once (far_tc_weight_split)."""
    lib = L.load()
    w = weight.detach()
    if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 2):
        raise L.FarError("tc_weight_split needs a contiguous fp32 CUDA [N,K] weight")
    N, K = w.shape
    nbytes = lib.far_tc_weight_split_bytes(N, K)
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=w.device)
    off = (-buf.data_ptr()) % 1024
    out = buf[off:off + nbytes]
    check(lib.far_tc_weight_split(ptr(w), K, N, K, ptr(out), nbytes, stream()), "far_tc_weight_split")
    return out


_ENC_KEYS = ("q_proj", "k_proj", "v_proj", "merge", "mlp0", "mlp2", "norm1_w", "norm1_b", "norm2_w", "norm2_b")


def loftr_encoder_layer(x, source, weights, nhead, engine=L.ENGINE_AUTO, eps1=1e-5, eps2=1e-5, presplit=None):
    """LoFTREncoderLayer.forward (transformer.py:44-67), masks None.  `weights`: dict with the layer's tensors
    q_proj,k_proj,v_proj,merge,mlp0,mlp2,norm1_w,norm1_b,norm2_w,norm2_b (CUDA fp32 contiguous, on x's device --
    checked: the C side reads raw pointers).  `presplit`: optional dict of tc_weight_split outputs for
    q_proj, kv ([k_proj; v_proj] stacked), merge, mlp0, mlp2 (cached by the module)."""
    lib = L.load()
    N, Lq, C = x.shape
    S = source.shape[1]
    x_, s_ = f32c(x), f32c(source)
    for k in _ENC_KEYS:
        t = weights[k]
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == x_.device):
            raise L.FarError(f"loftr_encoder_layer: weight '{k}' must be a contiguous fp32 CUDA tensor on {x_.device} "
                             f"(got {t.dtype}, {t.device}, contiguous={t.is_contiguous()})")
    out = torch.empty_like(x_)
    ps = presplit or {}
    w = L.EncoderLayerWeights(*[ptr(weights[k]) for k in _ENC_KEYS], float(eps1), float(eps2),
                              *[ptr(ps.get(k)) for k in ("q_proj", "kv", "merge", "mlp0", "mlp2")])
    nws = lib.far_loftr_encoder_layer_workspace_bytes(N, Lq, S, C, nhead)
    ws = _ws(nws, x.device)
    with _timed("far_loftr_encoder_layer"):
        check(lib.far_loftr_encoder_layer(ptr(x_), ptr(s_), ptr(out), N, Lq, S, C, nhead, byref(w), engine, ptr(ws),
                                        ws.numel(), stream()), "far_loftr_encoder_layer")
    return out


class _MatchHandle:
    """State between far_dual_softmax_match_select and _gather (the match count M is data dependent)."""
    __slots__ = ("ws", "nws", "cnt", "cnt_host", "event", "conf", "shape", "hw", "scales", "dev")


_SIDE_STREAMS = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def dual_softmax_match_begin(feat_c0, feat_c1, hw0_c, hw1_c, thr, border_rm, temperature, scale0, scale1,
                             return_conf_matrix=False, engine=L.ENGINE_AUTO, defer_readback=False):
    """First half of CoarseMatching (coarse_matching.py:86-193): launches the score / decision kernels and starts an
    asynchronous read-back of the match count on a side stream.  The caller may queue more GPU work (e.g. the FAR head
    trunk, which only needs the coarse features) before calling dual_softmax_match_end(): the host then learns M while
    the GPU is still busy, instead of draining the stream at the reference's torch.where sync point (:193)."""
    lib = L.load()
    N, Lq, C = feat_c0.shape
    S = feat_c1.shape[1]
    f0, f1 = f32c(feat_c0), f32c(feat_c1)
    dev = f0.device
    h = _MatchHandle()
    h.nws = lib.far_dual_softmax_match_workspace_bytes(N, Lq, S)
    h.ws = torch.empty(h.nws, dtype=torch.uint8, device=dev)  # private: must survive until _gather
    h.cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    h.conf = torch.empty((N, Lq, S), dtype=torch.float32, device=dev) if return_conf_matrix else None
    h.shape, h.hw, h.scales, h.dev = (N, Lq, S), (hw0_c, hw1_c), (float(scale0), float(scale1)), dev
    with _timed("far_dual_softmax_match_select"):
        check(lib.far_dual_softmax_match_select(ptr(f0), ptr(f1), N, Lq, S, C, float(temperature), float(thr),
                                              int(border_rm), hw0_c[0], hw0_c[1], hw1_c[0], hw1_c[1], ptr(h.conf),
                                              ptr(h.cnt), engine, ptr(h.ws), h.nws, stream()), "far_dual_softmax_match_select")
    h.event = None
    if not defer_readback:   # defer_readback: the caller captures this call in a CUDA graph and starts the copy after replay
        match_count_readback(h)
    return h


def match_count_readback(h):
    """Start the asynchronous device->host copy of the match count on the side stream (ordered after everything queued
    on the current stream so far, i.e. after the select kernels)."""
    dev = h.dev
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    ready = torch.cuda.Event()
    ready.record(cur)
    h.cnt_host = torch.empty(1, dtype=torch.int64, pin_memory=True)
    with torch.cuda.stream(side):
        side.wait_event(ready)
        h.cnt_host.copy_(h.cnt, non_blocking=True)
        h.event = torch.cuda.Event()
        h.event.record(side)
    h.cnt.record_stream(side)


def dual_softmax_match_end(h):
    """Second half (coarse_matching.py:193-263): waits for the match count only (not for the compute stream), then the
    order-preserving compaction."""
    lib = L.load()
    if h.event is None:
        match_count_readback(h)
    h.event.synchronize()   # the reference's sync point, but only on the count's own copy
    M = int(h.cnt_host[0])
    N, Lq, S = h.shape
    dev = h.dev
    b_ids = torch.empty(M, dtype=torch.int64, device=dev)
    i_ids = torch.empty(M, dtype=torch.int64, device=dev)
    j_ids = torch.empty(M, dtype=torch.int64, device=dev)
    mconf = torch.empty(M, dtype=torch.float32, device=dev)
    mk0 = torch.empty((M, 2), dtype=torch.float32, device=dev)
    mk1 = torch.empty((M, 2), dtype=torch.float32, device=dev)
    check(lib.far_dual_softmax_match_gather(N, Lq, h.hw[0][1], h.hw[1][1], h.scales[0], h.scales[1], M, ptr(b_ids),
                                            ptr(i_ids), ptr(j_ids), ptr(mconf), ptr(mk0), ptr(mk1), ptr(h.ws), h.nws,
                                            stream()), "far_dual_softmax_match_gather")
    out = {"b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids, "mconf": mconf, "mkpts0_c": mk0, "mkpts1_c": mk1}
    if h.conf is not None:
        out["conf_matrix"] = h.conf
    return out


def dual_softmax_match(feat_c0, feat_c1, hw0_c, hw1_c, thr, border_rm, temperature, scale0, scale1,
                       return_conf_matrix=False, engine=L.ENGINE_AUTO):
    """CoarseMatching.forward + get_coarse_match (coarse_matching.py:86-265), dual-softmax, eval.
    Returns dict(b_ids,i_ids,j_ids,mconf,mkpts0_c,mkpts1_c[,conf_matrix]).  One device->host read of M, as in the
    reference's torch.where (:193)."""
    return dual_softmax_match_end(dual_softmax_match_begin(feat_c0, feat_c1, hw0_c, hw1_c, thr, border_rm, temperature,
                                                           scale0, scale1, return_conf_matrix, engine))


def fine_preprocess(feat_f0, feat_f1, feat_c0, feat_c1, b_ids, i_ids, j_ids, W, stride, w0c, w1c, down_w, down_b,
                    merge_w, merge_b):
    """FinePreprocess.forward (fine_preprocess.py:29-59) with fine_concat_coarse_feat.  feat_f*: [N,Cf,Hf,Wf]
    (any of NCHW / channels_last), feat_c*: [N,L,Cc].  Returns two [M, W*W, Cf] tensors."""
    lib = L.load()
    dev = feat_f0.device
    n, Cf, Hf, Wf = feat_f0.shape
    M = int(b_ids.shape[0])
    out0 = torch.empty((M, W * W, Cf), dtype=torch.float32, device=dev)
    out1 = torch.empty((M, W * W, Cf), dtype=torch.float32, device=dev)
    if M == 0:
        return out0, out1
    if feat_f0.dtype != torch.float32:
        feat_f0, feat_f1 = feat_f0.float(), feat_f1.float()
    if feat_f0.stride() != feat_f1.stride() or not (feat_f0.stride(1) == 1 or feat_f0.stride(3) == 1):
        feat_f0, feat_f1 = feat_f0.contiguous(), feat_f1.contiguous()
    sn, sc, sh, sw = feat_f0.stride()
    c0, c1 = f32c(feat_c0), f32c(feat_c1)
    Cc = c0.shape[-1]
    nws = lib.far_fine_preprocess_workspace_bytes(M, W * W, Cf, Cc)
    ws = _ws(nws, dev)
    with _timed("far_fine_preprocess"):
        check(lib.far_fine_preprocess(ptr(feat_f0), ptr(feat_f1), sn, sc, sh, sw, Hf, Wf, Cf, ptr(c0), ptr(c1),
                                    c0.shape[1], c1.shape[1], Cc, ptr(b_ids), ptr(i_ids), ptr(j_ids), M, W, stride,
                                    w0c, w1c, ptr(f32c(down_w)), ptr(f32c(down_b)), ptr(f32c(merge_w)),
                                    ptr(f32c(merge_b)), ptr(out0), ptr(out1), ptr(ws), ws.numel(), stream()),
            "far_fine_preprocess")
    return out0, out1


def fine_match(feat_f0, feat_f1, mkpts1_c, offset_scale):
    """FineMatching.forward + get_fine_match (fine_matching.py:15-76).  Returns expec_f [M,3], mkpts1_f [M,2]."""
    lib = L.load()
    M, WW, C = feat_f0.shape
    dev = feat_f0.device
    expec = torch.empty((M, 3), dtype=torch.float32, device=dev)
    mk1f = torch.empty((M, 2), dtype=torch.float32, device=dev)
    if M == 0:
        return expec, mk1f
    with _timed("far_fine_match"):
        check(lib.far_fine_match(ptr(f32c(feat_f0)), ptr(f32c(feat_f1)), M, WW, C, ptr(f32c(mkpts1_c)),
                               float(offset_scale), ptr(expec), ptr(mk1f), stream()), "far_fine_match")
    return expec, mk1f


def eight_point(points1, points2, weights=None, counts=None):
    """run_8point (third_party/prior_ransac/cv_geometry.py:772-833).  points [P,N,2], weights [P,N] -> F [P,3,3]."""
    lib = L.load()
    P, N, _ = points1.shape
    if N < 8:
        raise AssertionError(points1.shape)  # the reference asserts N >= 8 (:787)
    p1, p2 = f32c(points1), f32c(points2)
    w = f32c(weights) if weights is not None else None
    c = counts.to(torch.int32).contiguous() if counts is not None else None
    F = torch.empty((P, 3, 3), dtype=torch.float32, device=p1.device)
    nws = lib.far_eight_point_workspace_bytes(P)
    ws = _ws(nws, p1.device)
    with _timed("far_eight_point"):
        check(lib.far_eight_point(ptr(p1), ptr(p2), ptr(w), ptr(c), P, N, ptr(F), ptr(ws), ws.numel(), stream()),
            "far_eight_point")
    return F


def essential_decompose(E):
    """decompose_essential_matrix (third_party/prior_ransac/essential.py:99-139): E [*,3,3] -> R1,R2 [*,3,3], t [*,3,1]."""
    lib = L.load()
    lead = E.shape[:-2]
    e = f32c(E).reshape(-1, 3, 3)
    P = e.shape[0]
    R1 = torch.empty_like(e)
    R2 = torch.empty_like(e)
    t = torch.empty((P, 3), dtype=torch.float32, device=e.device)
    check(lib.far_essential_decompose(ptr(e), P, ptr(R1), ptr(R2), ptr(t), stream()), "far_essential_decompose")
    return R1.reshape(*lead, 3, 3), R2.reshape(*lead, 3, 3), t.reshape(*lead, 3, 1)


def pose_from_matches(mkpts0, mkpts1, mconf, offsets, K0, K1):
    """Ragged per-pair weighted 8-point + cheirality-selected (R,t) (spvs_RT loop, supervision.py:184-233).
    offsets: int64 [N+1] segment boundaries (m_bids is sorted).  Returns E [N,3,3], Rt [N,3,4], n_pos [N] int32."""
    lib = L.load()
    N = offsets.shape[0] - 1
    dev = offsets.device
    E = torch.empty((N, 3, 3), dtype=torch.float32, device=dev)
    Rt = torch.empty((N, 3, 4), dtype=torch.float32, device=dev)
    npos = torch.empty(N, dtype=torch.int32, device=dev)
    nws = lib.far_eight_point_workspace_bytes(N)
    ws = _ws(nws, dev)
    with _timed("far_pose_from_matches"):
        check(lib.far_pose_from_matches(ptr(f32c(mkpts0)), ptr(f32c(mkpts1)), ptr(f32c(mconf)), ptr(offsets), N,
                                      ptr(f32c(K0)), ptr(f32c(K1)), ptr(E), ptr(Rt), ptr(npos), ptr(ws), ws.numel(),
                                      stream()), "far_pose_from_matches")
    return E, Rt, npos


def segment_offsets(m_bids, P):
    """offsets [P+1] int64 of the sorted batch ids m_bids [M] (replaces bincount + cumsum; one launch, no sync)."""
    lib = L.load()
    if not m_bids.is_cuda or m_bids.dtype != torch.int64:
        raise L.FarError("segment_offsets: m_bids must be a CUDA int64 tensor")
    m_bids = m_bids.contiguous()
    off = torch.empty(P + 1, dtype=torch.int64, device=m_bids.device)
    check(lib.far_segment_offsets(ptr(m_bids), int(m_bids.shape[0]), P, ptr(off), stream()), "far_segment_offsets")
    return off


def five_point(pts5):
    """Nister's 5-point solver (cv_geometry.py:861-1041) for explicit minimal samples: pts5 [S,5,4] fp64 calibrated
    (x1, y1, x2, y2) -> (E [S,10,3,3] fp64 with x2h^T E x1h = 0, zero padded; nsol [S] int32)."""
    lib = L.load()
    if not pts5.is_cuda:
        raise FarError("far_b200 ops need CUDA tensors")
    p = pts5.to(torch.float64).contiguous()
    S = p.shape[0]
    E = torch.empty((S, 10, 3, 3), dtype=torch.float64, device=p.device)
    ns = torch.empty(S, dtype=torch.int32, device=p.device)
    check(lib.far_five_point(ptr(p), S, ptr(E), ptr(ns), stream()), "far_five_point")
    return E, ns


def ransac_sample_models(mkpts0, mkpts1, offsets, K0, K1, prior_rt, bias_sigma_sq, H, seed, return_indices=False,
                         minimal_solver="8pt"):
    """RANSAC.sample + estimate_model_from_minsample (third_party/prior_ransac/ransac.py:161-175,250-253,358-367) for a
    ragged batch: H minimal models per pair from (prior-biased or uniform) samples drawn on the device with a
    counter-based Philox generator.  minimal_solver '8pt': the normalised 8-point on 8 draws; '5pt': the recipe's model
    type, Nister's 5-point on 5 draws disambiguated by a sixth.  Returns models [P,H,3,3] (and sample indices
    [P,H,8 or 6] int32)."""
    lib = L.load()
    P = offsets.shape[0] - 1
    dev = offsets.device
    M = int(mkpts0.shape[0])
    if minimal_solver not in ("8pt", "5pt"):
        raise FarError(f"unknown minimal solver {minimal_solver!r}")
    S = 8 if minimal_solver == "8pt" else 6
    models = torch.empty((P, H, 3, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((P, H, S), dtype=torch.int32, device=dev) if return_indices else None
    pr = f32c(prior_rt) if prior_rt is not None else None
    ws = _ws(lib.far_ransac_sample_models_workspace_bytes(M, P, H), dev)
    with _timed("far_ransac_sample_models"):
        check(lib.far_ransac_sample_models(ptr(f32c(mkpts0)), ptr(f32c(mkpts1)), ptr(offsets), M, P, ptr(f32c(K0)),
                                           ptr(f32c(K1)), ptr(pr), float(bias_sigma_sq), H, S,
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(models), ptr(idx), ptr(ws), ws.numel(),
                                           stream()), "far_ransac_sample_models")
    return (models, idx) if return_indices else models


def prior_ransac_score(mkpts0, mkpts1, offsets, K0, K1, models, prior_rt, pcl, prior_lambda, inl_th):
    """RANSAC.get_prior_estimate + verify + remove_bad_models (third_party/prior_ransac/ransac.py:203-231,256-292,
    303-308) for every pair of a ragged batch.  models [P,H,3,3]; prior_rt [P,3,4] or None; pcl [npcl,3].
    Returns scores [P,H], best_idx [P] int32, best_E [P,3,3], counts3 [P,3] int32, inlier_mask [M] uint8."""
    lib = L.load()
    P, H = models.shape[0], models.shape[1]
    dev = models.device
    M = int(mkpts0.shape[0])
    scores = torch.empty((P, H), dtype=torch.float32, device=dev)
    best = torch.empty(P, dtype=torch.int32, device=dev)
    bestE = torch.empty((P, 3, 3), dtype=torch.float32, device=dev)
    counts3 = torch.empty((P, 3), dtype=torch.int32, device=dev)
    mask = torch.zeros(max(M, 1), dtype=torch.uint8, device=dev)
    pr = f32c(prior_rt) if prior_rt is not None else None
    pc = f32c(pcl) if pcl is not None else None
    with _timed("far_prior_ransac_score"):
        check(lib.far_prior_ransac_score(ptr(f32c(mkpts0)), ptr(f32c(mkpts1)), ptr(offsets), P, ptr(f32c(K0)),
                                         ptr(f32c(K1)), ptr(f32c(models)), H, ptr(pr), ptr(pc),
                                         pc.shape[0] if pc is not None else 0, float(prior_lambda), float(inl_th),
                                         ptr(scores), ptr(best), ptr(bestE), ptr(counts3), ptr(mask), stream()),
              "far_prior_ransac_score")
    return scores, best, bestE, counts3, mask[:M]


def pose_from_essential(mkpts0, mkpts1, mask, offsets, K0, K1, E):
    """cv2.recoverPose-style candidate selection (metrics.py:164-170) for given essential matrices E [P,3,3], voting over
    the matches with mask != 0.  Returns Rt [P,3,4], n_pos [P] int32."""
    lib = L.load()
    P = E.shape[0]
    dev = E.device
    Rt = torch.empty((P, 3, 4), dtype=torch.float32, device=dev)
    npos = torch.empty(P, dtype=torch.int32, device=dev)
    ws = _ws(lib.far_pose_from_essential_workspace_bytes(P), dev)
    check(lib.far_pose_from_essential(ptr(f32c(mkpts0)), ptr(f32c(mkpts1)), ptr(mask) if mask is not None else None,
                                      ptr(offsets), P, ptr(f32c(K0)), ptr(f32c(K1)), ptr(f32c(E)), ptr(Rt), ptr(npos),
                                      ptr(ws), ws.numel(), stream()), "far_pose_from_essential")
    return Rt, npos


def emm_bilinear_attn(qkv1, qkv2, pos, num_heads, scale, engine=L.ENGINE_AUTO):
    """CrossAttention core (transformer.py:275-292): qkv1,qkv2 [B,N,3*C] (output of the shared qkv Linear),
    pos [1 or B, N, 6] -> fundamental_1, fundamental_2 [B,h,d+6,d+6]."""
    lib = L.load()
    B, N, C3 = qkv1.shape
    C = C3 // 3
    d = C // num_heads
    q1, q2, p_ = f32c(qkv1), f32c(qkv2), f32c(pos)
    dv = d + 6
    F1 = torch.empty((B, num_heads, dv, dv), dtype=torch.float32, device=q1.device)
    F2 = torch.empty_like(F1)
    nws = lib.far_emm_bilinear_attn_workspace_bytes(B, N, num_heads, d)
    ws = _ws(nws, q1.device)
    with _timed("far_emm_bilinear_attn"):
        check(lib.far_emm_bilinear_attn(ptr(q1), ptr(q2), ptr(p_), p_.shape[0], B, N, num_heads, d, float(scale), ptr(F1),
                                      ptr(F2), engine, ptr(ws), ws.numel(), stream()), "far_emm_bilinear_attn")
    return F1, F2


_CORR_GRID = {}


def corr_volume_warp(vol0, vol1):
    """CorrelationVolumeWarping.forward (mapfree_6dreg/lib/models/regression/aggregator.py:42-116; POSITION_ENCODER +
    MAX_SCORE_CHANNEL).  vol0, vol1 [B,32,H,W] (NCHW or channels_last) -> [B, 67, H, W] = cat[vol0, warped vol1,
    soft position (2), max score (1)]."""
    lib = L.load()
    B, D, H, W = vol0.shape
    if vol0.shape != vol1.shape:
        raise AssertionError('Feature volumes shape must match')
    if vol0.dtype != torch.float32 or vol1.dtype != torch.float32:
        vol0, vol1 = vol0.float(), vol1.float()
    if vol0.stride() != vol1.stride() or not (vol0.stride(1) == 1 or vol0.is_contiguous()):
        vol0, vol1 = vol0.contiguous(), vol1.contiguous()
    N = H * W
    if vol0.is_contiguous():
        sb, sc, sp = D * N, N, 1
    else:  # channels_last: element (b, c, y, x) at b*sb + (y*W + x)*D + c
        if vol0.stride() != (N * D, 1, W * D, D):
            vol0, vol1 = vol0.contiguous(), vol1.contiguous()
            sb, sc, sp = D * N, N, 1
        else:
            sb, sc, sp = N * D, 1, D
    dev = vol0.device
    key = (H, W, str(dev))
    grid = _CORR_GRID.get(key)
    if grid is None:  # aggregator.py:84-88 (torch.meshgrid default indexing = 'ij')
        uu, vv = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing='ij')
        grid = torch.stack([uu, vv], dim=0).reshape(2, N).contiguous().to(dev)
        _CORR_GRID[key] = grid
    out = torch.empty((B, 2 * D + 3, H, W), dtype=torch.float32, device=dev)
    ws = _ws(lib.far_corr_volume_warp_workspace_bytes(B, N), dev)
    with _timed("far_corr_volume_warp"):
        check(lib.far_corr_volume_warp(ptr(vol0), ptr(vol1), sb, sc, sp, ptr(grid), B, N, D, ptr(out), ptr(ws),
                                       ws.numel(), stream()), "far_corr_volume_warp")
    return out


def softmax_attention(qkv, num_heads, scale):
    """timm Attention core (vision_transformer.py:250-257): qkv [B,N,3*C] -> [B,N,C]."""
    lib = L.load()
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // num_heads
    q = f32c(qkv)
    out = torch.empty((B, N, C), dtype=torch.float32, device=q.device)
    nws = lib.far_softmax_attention_workspace_bytes(B, N, num_heads, d)
    ws = _ws(nws, q.device)
    with _timed("far_softmax_attention"):
        check(lib.far_softmax_attention(ptr(q), B, N, num_heads, d, float(scale), ptr(out), ptr(ws), ws.numel(), stream()),
            "far_softmax_attention")
    return out


def pose_blend_mp3d(pred, solver, wt, mean9, std9, scale_8pt=True):
    """Gated fusion epilogue of forward_emm (transformer.py:436-469)."""
    lib = L.load()
    B = pred.shape[0]
    s = f32c(solver)
    out = torch.empty((B, 9), dtype=torch.float32, device=pred.device)
    check(lib.far_pose_blend_mp3d(ptr(f32c(pred)), ptr(s), s.shape[1], ptr(f32c(wt)), ptr(f32c(mean9)),
                                  ptr(f32c(std9)), int(scale_8pt), ptr(out), B, stream()), "far_pose_blend_mp3d")
    return out
