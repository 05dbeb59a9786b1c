#!/usr/bin/env python
"""The tcgen05 GEMM kernels in the regime of the real step: launches back to back (GPU at its power cap, no idle gaps),
inputs rotating over buffers larger than L2.  One subprocess per FAR_TC_TS mode (0 = SS kernel, 1 = A via TMEM,
2 = CTA pairs).  usage: python benchmarks/bench_gemm_hot.py [modes...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, torch
sys.path.insert(0, %r)
from far_b200 import ops
from far_b200._lib import ENGINE_TCGEN05, ACT_NONE, ACT_ELU1, ACT_RELU
M = 307200
res = {}
for name, (N, K, act, two) in {"256x256": (256, 256, ACT_NONE, False), "512x256_elu": (512, 256, ACT_ELU1, False),
                                 "512x512_2seg_relu": (512, 512, ACT_RELU, True), "256x512": (256, 512, ACT_NONE, False)}.items():
    xs = [torch.randn(M, K // 2 if two else K, device="cuda") for _ in range(3)]
    x2s = [torch.randn(M, K // 2, device="cuda") for _ in range(3)] if two else [None] * 3
    w = torch.randn(N, K, device="cuda") * 0.05
    ys = [torch.empty(M, N, device="cuda") for _ in range(2)]
    def run(i):
        ops.linear(xs[i %% 3], w, None, act, x2=x2s[i %% 3], engine=ENGINE_TCGEN05, out=ys[i %% 2])
    for i in range(20): run(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(60): run(i)
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 60 * 1e3
    res[name] = {"us": round(us, 1), "TF_alg": round(2.0 * M * N * K / us / 1e6, 1)}
    del xs, x2s, ys
print(json.dumps(res))
''' % ROOT
for mode in sys.argv[1:] or ("0", "1", "2"):
    env = dict(os.environ, FAR_TC_TS=mode)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(f"FAR_TC_TS={mode}", out.stdout.strip(), out.stderr.strip()[-300:], flush=True)
