#!/usr/bin/env python
"""Backbone (ResNetFPN_8_2, cuDNN convs) timing: plain module graph vs the eval-time fused path, 64 images 480x640."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from far_b200.loftr.backbone import ResNetFPN_8_2  # noqa: E402
from far_b200 import synth  # noqa: E402

torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
cfg = {'initial_dim': 128, 'block_dims': [128, 196, 256]}
x = torch.rand(64, 1, 480, 640, device='cuda')
m = ResNetFPN_8_2(cfg)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), 1234))
m = m.cuda().eval()
ref = None
for fused in (False, True):
    m.fused_eval = fused
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
    print(json.dumps({"fused_eval": fused, "ms": s.elapsed_time(e) / 5, "max_diff_vs_plain": err,
                      "feat_f_stride": list(f.stride())}), flush=True)
    del c, f
    torch.cuda.empty_cache()
