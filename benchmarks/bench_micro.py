#!/usr/bin/env python
"""Kernel micro-benchmarks with roofline fractions (BASELINE.json configs[4]: dual-softmax + 8-point solver,
4096 pairs x 2048 correspondences) and per-op timings of the hot-path kernels at the 640x480 shapes.

    python benchmarks/bench_micro.py [--pairs 4096] [--corr 2048] [--quick]

One JSON line per kernel: algorithmic bytes/FLOPs / CUDA-event time (median of timed iterations after
warm-up, L2 flushed between iterations by writing a 256 MB buffer) against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from far_b200 import ops, synth  # noqa: E402
from far_b200._lib import ENGINE_SIMT, ENGINE_TCGEN05, ACT_NONE  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def timeit(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)  # > L2
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--corr", type=int, default=2048)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    hbm, tf, which = peaks()
    flush = torch.zeros(64 * 1024 * 1024, device=dev)  # 256 MB > 126 MB L2
    out = []

    def emit(name, ms, bytes_=None, flops=None, **kw):
        r = {"kernel": name, "ms": ms, **kw}
        if bytes_ is not None:
            r.update({"bound": "hbm", "achieved_GBs": bytes_ / ms / 1e6, "peak_GBs": hbm, "frac": bytes_ / ms / 1e6 / hbm})
        if flops is not None:
            r.update({"bound": "tensor", "achieved_TFs": flops / ms / 1e9, "peak_TFs": tf, "frac": flops / ms / 1e9 / tf})
        r["peak_source"] = which
        print(json.dumps(r), flush=True)
        out.append(r)

    # ---- config 5 (i): weighted 8-point + decomposition, P pairs x N correspondences
    P, N = (512, a.corr) if a.quick else (a.pairs, a.corr)
    p1, p2, w, _, _ = synth.two_view_geometry(min(P, 256), N, seed=5)
    rep = (P + p1.shape[0] - 1) // p1.shape[0]
    p1, p2, w = [t.repeat(rep, *([1] * (t.dim() - 1)))[:P].contiguous().to(dev) for t in (p1, p2, w)]
    ms = timeit(lambda: ops.eight_point(p1, p2, w), flush=flush)
    emit("far_eight_point (accumulate + solve)", ms, bytes_=20.0 * N * P + 36 * P, pairs=P, corr=N,
         pairs_per_s=P / ms * 1e3)
    F = ops.eight_point(p1, p2, w)
    ms = timeit(lambda: ops.essential_decompose(F), flush=flush)
    emit("far_essential_decompose", ms, bytes_=(36 + 84.0) * P, pairs=P)

    # ---- config 5 (ii): dual-softmax + mutual-NN match on L = S = 2048 tokens, chunks of pairs
    L, C, chunk = 2048, 256, 16 if a.quick else 64
    g = np.random.default_rng(1)
    f0 = torch.from_numpy(g.standard_normal((chunk, L, C)).astype(np.float32)).to(dev)
    f1 = torch.from_numpy(g.standard_normal((chunk, L, C)).astype(np.float32)).to(dev)
    ms = timeit(lambda: ops.dual_softmax_match(f0, f1, (32, 64), (32, 64), 0.0, 0, 0.1, 8.0, 8.0), iters=5, flush=flush)
    emit("far_dual_softmax_match (select+gather, fused from features)", ms, flops=2 * 2.0 * L * L * C * chunk,
         pairs=chunk, tokens=L, pairs_per_s=chunk / ms * 1e3)

    # ---- hot-path kernels at the 640x480 shapes (8 pairs)
    n = 2 if a.quick else 8
    x = torch.randn(n * 4800, 256, device=dev)
    wgt = torch.randn(256, 256, device=dev) / 16
    for eng, nm in ((ENGINE_SIMT, "cuda-core fp32"), (ENGINE_TCGEN05, "tcgen05 3xTF32")):
        try:
            ms = timeit(lambda: ops.linear(x, wgt, None, ACT_NONE, engine=eng), flush=flush)
            emit(f"far_linear 256x256 [{nm}]", ms, flops=2.0 * x.shape[0] * 256 * 256, rows=x.shape[0])
        except Exception as ex:  # engine unsupported on this shape
            print(json.dumps({"kernel": f"far_linear [{nm}]", "error": str(ex)}))
    q = torch.randn(n, 4800, 8, 32, device=dev)
    ms = timeit(lambda: ops.linear_attention(q, q, q), flush=flush)
    emit("far_linear_attention L=S=4800", ms, bytes_=4.0 * n * 4800 * 256 * 4, pairs=n)
    qkv = torch.randn(n, 4800, 768, device=dev) / 4
    pos = torch.rand(1, 4800, 6, device=dev)
    ms = timeit(lambda: ops.emm_bilinear_attn(qkv, qkv, pos, 4, 0.125), iters=3, warm=1, flush=flush)
    emit("far_emm_bilinear_attn N=4800 h=4 d=64 (both directions)", ms,
         flops=2 * n * 4 * (2 * 2.0 * 4800 * 4800 * 64 + 2.0 * 4800 * 4800 * 70), pairs=n)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_micro.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
