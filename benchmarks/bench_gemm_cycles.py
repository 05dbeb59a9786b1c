#!/usr/bin/env python
"""Cycle-level utilisation of tc_gemm_pair_kernel without perturbing its loops (FAR_TC_DBG bit 1024: two clock reads per
CTA): tensor-pipe floor = tiles x k-blocks x 768 cycles (12 M256xN128xK8 tf32 MMAs of 64 cycles per 32-wide k-block)
against the measured cycles of the slowest CTA, and the SM clock implied by the CUDA-event time."""
import ctypes
import os
import sys

os.environ["FAR_TC_DBG"] = str(int(os.environ.get("FAR_TC_DBG", "0")) | 1024)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from far_b200 import ops, _lib
from far_b200._lib import ENGINE_TCGEN05, ACT_NONE, ACT_ELU1, ACT_RELU

lib = _lib.load()
M = int(os.environ.get("M", 153600))
shapes = {"256x256": (256, 256, ACT_NONE, False), "256x256_elu": (256, 256, ACT_ELU1, False), "512x256": (512, 256, ACT_NONE, False),
          "512x512_2seg_relu": (512, 512, ACT_RELU, True), "256x512": (256, 512, ACT_NONE, False)}
ONLY = os.environ.get("ONLY")
for name, (N, K, act, two) in shapes.items():
    if ONLY and name != ONLY:
        continue
    if two:
        x = torch.randn(M, K // 2, device="cuda"); x2 = torch.randn(M, K // 2, device="cuda")
    else:
        x = torch.randn(M, K, device="cuda"); x2 = None
    w = torch.randn(N, K, device="cuda") * 0.05
    y = torch.empty(M, N, device="cuda")
    for _ in range(int(os.environ.get("WARM", 10))):
        ops.linear(x, w, None, act, x2=x2, engine=ENGINE_TCGEN05, out=y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.linear(x, w, None, act, x2=x2, engine=ENGINE_TCGEN05, out=y); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    buf = (ctypes.c_ulonglong * 320)()
    assert lib.far_tc_debug_counters(3, buf) == 0
    c = np.array(list(buf), dtype=np.float64).reshape(160, 2)
    c = c[c[:, 1] > 0]
    kb = (K + 31) // 32
    floor = c[:, 1] * kb * 768
    worst = int(np.argmax(c[:, 0]))
    us = float(np.median(ts))
    print(f"{name}: {us:.1f} us  CTAs {len(c)}  cycles max {c[:, 0].max():.0f} (tiles {c[worst, 1]:.0f}) mean {c[:, 0].mean():.0f}  "
          f"tensor floor of the slowest CTA {floor[worst]:.0f} -> pipe utilisation {floor[worst] / c[worst, 0]:.3f}  "
          f"cycles/tile {c[worst, 0] / c[worst, 1]:.0f} (floor {kb * 768})  implied SM clock {c[:, 0].max() / us / 1e3:.2f} GHz")
