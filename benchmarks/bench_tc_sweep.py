#!/usr/bin/env python
"""tcgen05 GEMM engine cost model: time far_linear(engine=tcgen05) over (tiles per CTA, k-blocks) to separate the fixed
launch cost, the per-tile cost and the per-k-block cost.  CUDA events, median of 20, not under a profiler."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from far_b200 import ops  # noqa: E402
from far_b200._lib import ENGINE_TCGEN05, ENGINE_SIMT, ACT_NONE  # noqa: E402


def t(fn, it=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(it):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)) * 1e3  # us


for tiles_per_cta in (1, 4, 16):
    for N in (128, 256):
        for K in (32, 256, 1024):
            M = 128 * 148 * tiles_per_cta // (N // 128)
            x = torch.randn(M, K, device="cuda")
            w = torch.randn(N, K, device="cuda") / K ** 0.5
            us = t(lambda: ops.linear(x, w, None, ACT_NONE, engine=ENGINE_TCGEN05))
            us2 = t(lambda: ops.linear(x, w, None, ACT_NONE, engine=ENGINE_SIMT))
            print(json.dumps({"M": M, "N": N, "K": K, "tiles_per_cta": tiles_per_cta, "kblocks": K // 32,
                              "tc_us": round(us, 1), "simt_us": round(us2, 1),
                              "tc_TFs": round(2.0 * M * N * K / us / 1e6, 1)}), flush=True)
