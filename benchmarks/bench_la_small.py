#!/usr/bin/env python
"""Fine-window linear attention (la_small_allheads_kernel): N windows x 25 tokens x 8 heads x 16 dims, back to back.
usage: python benchmarks/bench_la_small.py [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from far_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 41800
q, k, v = (torch.rand(N, 25, 8, 16, device="cuda") + 0.1 for _ in range(3))   # already-mapped (positive) features, as in the fused layer
for _ in range(5):
    ops.linear_attention(q, k, v, feature_map_applied=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(40):
    ops.linear_attention(q, k, v, feature_map_applied=True)
e.record(); torch.cuda.synchronize()
us = s.elapsed_time(e) / 40 * 1e3
gb = 4 * N * 25 * 128 * 4 / 1e9
print(f"la_small: N={N}: {us:.1f} us per call, {gb / us * 1e6:.0f} GB/s algorithmic ({gb:.2f} GB)")
