#!/usr/bin/env python
"""One LoFTREncoderLayer call at the headline shape (N pairs-sides x 4800 tokens x 256) with the library's per-kernel
profiler on: fused schedule (engine 0) vs kernel-per-op tensor-core schedule (engine 3) and the difference between the
two (parity against the oracle lives in tests/test_gpu_tcgen05.py::test_fused_encoder_layer_schedule).  usage: python benchmarks/bench_layer.py [N]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from far_b200 import ops, _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L, C, H = 4800, 256, 8
dev = "cuda"
g = torch.Generator().manual_seed(3)
x = (torch.randn(N, L, C, generator=g) * 1.5).to(dev)
src = (torch.randn(N, L, C, generator=g) * 1.5).to(dev)
w = {}
for k, shp in (("q_proj", (C, C)), ("k_proj", (C, C)), ("v_proj", (C, C)), ("merge", (C, C)), ("mlp0", (2 * C, 2 * C)),
               ("mlp2", (C, 2 * C))):
    bound = (6.0 / (shp[0] + shp[1])) ** 0.5
    w[k] = ((torch.rand(*shp, generator=g) * 2 - 1) * bound).to(dev)
for k in ("norm1_w", "norm2_w"):
    w[k] = (1 + 0.1 * torch.randn(C, generator=g)).to(dev)
for k in ("norm1_b", "norm2_b"):
    w[k] = (0.1 * torch.randn(C, generator=g)).to(dev)

outs = {}
for name, eng in (("fused", 0), ("per_op", 3)):
    for _ in range(3):
        out = ops.loftr_encoder_layer(x, src, w, H, eng)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        out = ops.loftr_encoder_layer(x, src, w, H, eng)
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    outs[name] = out
    line = {"schedule": name, "N": N, "ms_per_call": e0.elapsed_time(e1) / iters,
            "kernels": {k: {"us_avg": 1e3 * v["ms"] / v["launches"], "launches_per_call": v["launches"] / iters,
                            "TFLOPs": v["flops"] / v["ms"] / 1e9, "GBs": v["bytes"] / v["ms"] / 1e6}
                        for k, v in prof.items()}}
    print(json.dumps(line))
d = (outs["fused"] - outs["per_op"]).abs().max().item()
print(json.dumps({"fused_vs_per_op_max_abs_diff": d, "finite": bool(torch.isfinite(outs["fused"]).all())}))
