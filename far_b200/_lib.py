"""ctypes binding of libfar_sm100.so (the C ABI declared in include/far_sm100.h).

There is no CPU fallback and no PyTorch-eager fallback: if the shared library is missing, or a tensor is not a
CUDA fp32 tensor, the op raises.  PyTorch is used for device memory, streams and torch.distributed only.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfar_sm100.so")

FAR_OK, FAR_ERR_ARG, FAR_ERR_CUDA, FAR_ERR_WORKSPACE = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_GELU, ACT_ELU1, ACT_SIGMOID = 0, 1, 2, 3, 4
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
_ERR = {1: "FAR_ERR_ARG (shape/alignment/null precondition)", 2: "FAR_ERR_CUDA (kernel launch failed)",
        3: "FAR_ERR_WORKSPACE (workspace too small)"}


class FarError(RuntimeError):
    pass


class EncoderLayerWeights(Structure):
    _fields_ = [(n, c_void_p) for n in ("wq", "wk", "wv", "wmerge", "wmlp0", "wmlp2", "g1", "b1", "g2", "b2")] + \
        [("eps1", c_float), ("eps2", c_float)] + \
        [(n, c_void_p) for n in ("ps_wq", "ps_wkv", "ps_wmerge", "ps_wmlp0", "ps_wmlp2")]


_P = c_void_p
_SIGS = {
    "far_abi_version": (c_int, []),
    "far_launch_count": (ctypes.c_ulonglong, []),
    "far_tc_set_cross16": (c_int, [c_int]),
    "far_tc_debug_counters": (c_int, [c_int, _P]),
    "far_tc_weight_split_bytes": (c_size_t, [c_int, c_int]),
    "far_tc_weight_split": (c_int, [_P, c_int, c_int, c_int, _P, c_size_t, _P]),
    "far_linear_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "far_linear": (c_int, [_P, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int,
                           c_int, _P, c_size_t, _P]),
    "far_layernorm": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "far_layernorm_pre": (c_int, [_P, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "far_upsample2x_add_nhwc": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "far_scale_shift_act_nhwc": (c_int, [_P, _P, _P, c_longlong, c_int, c_float, _P]),
    "far_scale_shift_act_nhwc_out": (c_int, [_P, _P, _P, _P, c_longlong, c_int, c_float, _P]),
    "far_vit_preprocess": (c_int, [_P, _P, c_longlong, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), _P]),
    "far_stem_conv_workspace_bytes": (c_size_t, [c_int]),
    "far_stem_conv7x7s2_relu_nhwc": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "far_pos_encode_flatten": (c_int, [_P, c_longlong, c_longlong, c_longlong, c_longlong, _P, _P, c_int, c_int, c_int,
                                       c_int, _P]),
    "far_linear_attention_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "far_linear_attention": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_float, c_int, _P, c_size_t, _P]),
    "far_loftr_encoder_layer_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "far_loftr_encoder_layer": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, POINTER(EncoderLayerWeights),
                                        c_int, _P, c_size_t, _P]),
    "far_dual_softmax_match_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "far_dual_softmax_match_select": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int,
                                              c_int, c_int, c_int, _P, _P, c_int, _P, c_size_t, _P]),
    "far_dual_softmax_match_gather": (c_int, [c_int, c_int, c_int, c_int, c_float, c_float, c_longlong, _P, _P, _P, _P,
                                              _P, _P, _P, c_size_t, _P]),
    "far_fine_preprocess_workspace_bytes": (c_size_t, [c_longlong, c_int, c_int, c_int]),
    "far_fine_preprocess": (c_int, [_P, _P, c_longlong, c_longlong, c_longlong, c_longlong, c_int, c_int, c_int, _P, _P,
                                    c_int, c_int, c_int, _P, _P, _P, c_longlong, c_int, c_int, c_int, c_int, _P, _P, _P,
                                    _P, _P, _P, _P, c_size_t, _P]),
    "far_fine_match": (c_int, [_P, _P, c_longlong, c_int, c_int, _P, c_float, _P, _P, _P]),
    "far_eight_point_workspace_bytes": (c_size_t, [c_int]),
    "far_eight_point": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P, c_size_t, _P]),
    "far_essential_decompose": (c_int, [_P, c_int, _P, _P, _P, _P]),
    "far_pose_from_matches": (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "far_prior_ransac_score": (c_int, [_P, _P, _P, c_int, _P, _P, _P, c_int, _P, _P, c_int, c_float, c_float, _P, _P, _P,
                                       _P, _P, _P]),
    "far_segment_offsets": (c_int, [_P, c_longlong, c_int, _P, _P]),
    "far_five_point": (c_int, [_P, c_int, _P, _P, _P]),
    "far_ransac_sample_models_workspace_bytes": (c_size_t, [c_longlong, c_int, c_int]),
    "far_ransac_sample_models": (c_int, [_P, _P, _P, c_longlong, c_int, _P, _P, _P, c_float, c_int, c_int,
                                         ctypes.c_ulonglong, _P, _P, _P, c_size_t, _P]),
    "far_pose_from_essential_workspace_bytes": (c_size_t, [c_int]),
    "far_pose_from_essential": (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "far_emm_bilinear_attn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "far_emm_bilinear_attn": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, c_int, _P,
                                      c_size_t, _P]),
    "far_corr_volume_warp_workspace_bytes": (c_size_t, [c_int, c_int]),
    "far_corr_volume_warp": (c_int, [_P, _P, c_longlong, c_longlong, c_longlong, _P, c_int, c_int, c_int, _P, _P,
                                     c_size_t, _P]),
    "far_softmax_attention_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "far_softmax_attention": (c_int, [_P, c_int, c_int, c_int, c_int, c_float, _P, _P, c_size_t, _P]),
    "far_pose_blend_mp3d": (c_int, [_P, _P, c_int, _P, _P, _P, c_int, _P, c_int, _P]),
    "far_profile_num_ids": (c_int, []),
    "far_profile_name": (ctypes.c_char_p, [c_int]),
    "far_profile_enable": (c_int, [c_int]),
    "far_profile_read": (c_int, [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong),
                                 ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
}

_lib = None


def exported_symbols():
    """Names declared in include/far_sm100.h (used by the CPU-side symbol test)."""
    return sorted(_SIGS)


def load():
    """dlopen the library (once) and attach signatures.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FarError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(far_b200 has no CPU / eager fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_profiling = False


def profile_enable(on):
    global _profiling
    check(load().far_profile_enable(int(on)), "far_profile_enable")
    _profiling = bool(on)


def profiling_enabled():
    return _profiling


def profile_read():
    """{kernel class: dict(ms, launches, flops, bytes)} from the library's per-launch CUDA events (far_sm100.h)."""
    lib = load()
    out = {}
    for i in range(lib.far_profile_num_ids()):
        ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        n = ctypes.c_ulonglong()
        check(lib.far_profile_read(i, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl), ctypes.byref(by)),
              "far_profile_read")
        if n.value:
            out[lib.far_profile_name(i).decode()] = {"ms": ms.value, "launches": n.value, "flops": fl.value,
                                                     "bytes": by.value}
    return out


def check(rc, what):
    if rc != FAR_OK:
        raise FarError(f"{what} failed: {_ERR.get(rc, rc)}")


def ptr(t):
    """Device pointer of a CUDA fp32/int64/int32 tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise FarError("far_b200 ops need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def f32c(t):
    """Contiguous fp32 CUDA view/copy."""
    if not t.is_cuda:
        raise FarError("far_b200 ops need CUDA tensors (there is no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def stream():
    return torch.cuda.current_stream().cuda_stream


class _Workspace:
    """Grow-only scratch buffer per (device, CUDA stream).  Ops issued on one stream run in order, so one buffer per
    stream is enough; two streams (a side-stream prefetch, DataParallel threads with their own streams) never alias.
    Ops whose workspace must survive until a later call (match select -> gather) take a private allocation instead."""

    def __init__(self):
        self.buf = {}

    def get(self, nbytes, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream(device).cuda_stream)
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


workspace = _Workspace()
