#!/usr/bin/env python
"""tc_gemm_kernel at the encoder-layer shapes, with the FAR_TC_DBG ablations (2 = no epilogue body, 4 = no MMAs,
6 = neither: pure TMA + converter pipeline) to see which stage bounds the kernel.
usage: python benchmarks/bench_gemm.py            (spawns one subprocess per ablation)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, torch, numpy as np
sys.path.insert(0, %r)
from far_b200 import ops, _lib
from far_b200._lib import ENGINE_TCGEN05, ACT_NONE, ACT_ELU1, ACT_RELU
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
def t(fn, it=15):
    for _ in range(3): fn()
    ts = []
    for _ in range(it):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts)) * 1e3
M = 153600
res = {}
for name, (N, K, act, two) in {"256x256": (256, 256, ACT_NONE, False), "256x256_elu": (256, 256, ACT_ELU1, False),
                                 "512x256": (512, 256, ACT_NONE, False), "512x512_2seg_relu": (512, 512, ACT_RELU, True),
                                 "256x512": (256, 512, ACT_NONE, False)}.items():
    if two:
        x = torch.randn(M, K // 2, device="cuda"); x2 = torch.randn(M, K // 2, device="cuda")
    else:
        x = torch.randn(M, K, device="cuda"); x2 = None
    w = torch.randn(N, K, device="cuda") * 0.05
    y = torch.empty(M, N, device="cuda")
    us = t(lambda: ops.linear(x, w, None, act, x2=x2, engine=ENGINE_TCGEN05, out=y))
    res[name] = {"us": round(us, 1), "TF_alg": round(2.0 * M * N * K / us / 1e6, 1),
                 "GBs": round(4.0 * (M * K + M * N + N * K) / us / 1e3, 0)}
print(json.dumps(res))
''' % ROOT
for dbg in sys.argv[1:] or ("0", "2", "4", "6"):
    env = dict(os.environ, FAR_TC_DBG=dbg)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(f"FAR_TC_DBG={dbg}", out.stdout.strip(), out.stderr.strip()[-300:])
