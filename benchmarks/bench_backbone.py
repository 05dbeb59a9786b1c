#!/usr/bin/env python
"""Backbone (ResNetFPN_8_2 on cuDNN) timing variants: memory format x BN folding, 64 images 480x640."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from far_b200.loftr.backbone import ResNetFPN_8_2, fold_batchnorm  # noqa: E402
from far_b200 import synth  # noqa: E402

torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
cfg = {'initial_dim': 128, 'block_dims': [128, 196, 256]}
x = torch.rand(64, 1, 480, 640, device='cuda')
ref = None
for fmt in ("channels_last", "contiguous"):
    for fold in (False, True):
        m = ResNetFPN_8_2(cfg)
        m.load_state_dict(synth.synth_state_dict(m.state_dict(), 1234))
        m = m.cuda().eval()
        m.memory_format = torch.channels_last if fmt == "channels_last" else torch.contiguous_format
        if fmt == "channels_last":
            m = m.to(memory_format=torch.channels_last)
        if fold:
            m = fold_batchnorm(m)
        with torch.no_grad():
            for _ in range(3):
                c, f = m(x)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                c, f = m(x)
            e.record()
            torch.cuda.synchronize()
        if ref is None:
            ref = (c.float().clone(), f.float().clone())
        err = max((c - ref[0]).abs().max().item(), (f - ref[1]).abs().max().item())
        print(json.dumps({"format": fmt, "bn_folded": fold, "ms": s.elapsed_time(e) / 5, "max_diff_vs_first": err,
                          "feat_f_stride": list(f.stride())}), flush=True)
