#!/usr/bin/env python
"""The [32 x 35840] x [512] FAR encoder / gate MLP GEMM (73 MB of weights) over rotating weight buffers (> L2).
FAR_SKINNY=0: 128x128 tile kernel + split-K (113.6 us on B200); 1 / unset: 32-row tile kernel (56.1 us).  A row-streaming
variant (a warp per two weight rows, 2 KB runs along K, activations in shared memory) measured 118.8 us and was dropped."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from far_b200 import ops
from far_b200._lib import ACT_RELU

M, N, K = 32, 512, 35840
x = torch.randn(M, K, device="cuda")
ws = [torch.randn(N, K, device="cuda") * 0.01 for _ in range(4)]
b = torch.randn(N, device="cuda")
for i in range(8):
    ops.linear(x, ws[i % 4], b, ACT_RELU)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(40):
    ops.linear(x, ws[i % 4], b, ACT_RELU)
e.record(); torch.cuda.synchronize()
us = s.elapsed_time(e) / 40 * 1e3
print(f"FAR_SKINNY={os.environ.get('FAR_SKINNY', '1')}: {us:.1f} us per call, {N * K * 4 / us / 1e3:.0f} GB/s of weights")
