#!/usr/bin/env python
"""Per-role wait accounting of tc_gemm_pair_kernel (default) / tc_gemm_ts_kernel (FAR_TC_TS=1) (FAR_TC_DBG bit 256, CTA 0): which pipeline barrier each warp role
waits on, as cycles per k-block.  usage: FAR_TC_DBG=256 python benchmarks/bench_gemm_roles.py"""
import ctypes
import os
import sys

# bit 256: per-tile waits only (cheap); HEAVY=1 adds bit 2048: per-k-block sites and the timeline (perturbs the MMA warp)
os.environ["FAR_TC_DBG"] = str(int(os.environ.get("FAR_TC_DBG", "0")) | 256 | (2048 if os.environ.get("HEAVY") else 0))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from far_b200 import ops, _lib
from far_b200._lib import ENGINE_TCGEN05, ACT_NONE, ACT_ELU1, ACT_RELU

lib = _lib.load()
KERNEL = 1 if os.environ.get("FAR_TC_TS", "2") == "2" else 0   # which kernel's counters: 1 = CTA-pair, 0 = one-CTA TS
M = int(os.environ.get("M", 153600))
WARM = int(os.environ.get("WARM", 10))
names = ["prod wait_empty", "prod total", "mma wait_main", "mma wait_cross", "mma wait_conv", "mma total",
         "conv wait_full", "conv wait_afree", "conv total", "epi wait_tfull", "epi total", "epi tfull->cross", "(tiles)", "(k-blocks)", "prod prefetch", "prod tma issue"]
for name, (N, K, act, two) in {"256x256": (256, 256, ACT_NONE, False), "256x256_elu": (256, 256, ACT_ELU1, False),
                               "512x512_2seg_relu": (512, 512, ACT_RELU, True), "256x512": (256, 512, ACT_NONE, False)}.items():
    if two:
        x = torch.randn(M, K // 2, device="cuda"); x2 = torch.randn(M, K // 2, device="cuda")
    else:
        x = torch.randn(M, K, device="cuda"); x2 = None
    w = torch.randn(N, K, device="cuda") * 0.05
    y = torch.empty(M, N, device="cuda")
    for _ in range(WARM):
        ops.linear(x, w, None, act, x2=x2, engine=ENGINE_TCGEN05, out=y)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.linear(x, w, None, act, x2=x2, engine=ENGINE_TCGEN05, out=y); e.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    assert lib.far_tc_debug_counters(KERNEL, buf) == 0
    c = list(buf)
    tiles, kb = max(c[12], 1), max(c[13], 1)
    print(f"{name}: {s.elapsed_time(e) * 1e3:.1f} us, CTA0: {tiles} tiles x {kb} k-blocks, mma total {c[5]} clk "
          f"= {c[5] / (tiles * kb):.0f} clk/k-block")
    for i, n in enumerate(names):
        print(f"    {n:18s} {c[i]:10d}  {c[i] / (tiles * kb):8.1f} /k-block  {c[i] / tiles:9.0f} /tile")
    if KERNEL == 1 and os.environ.get("FAR_TRACE"):
        tr = (ctypes.c_ulonglong * 256)()
        assert lib.far_tc_debug_counters(2, tr) == 0
        t = list(tr)
        t0 = t[0]
        rel = lambda v: (v - t0) if v else None
        print(f"    timeline of tile 4 (cycles from the MMA warp's tile start); epilogue of tile 3: tfull {rel(t[164])} "
              f"cross back {rel(t[165])} main back {rel(t[166])} end {rel(t[167])}")
        print(f"      cross_empty seen by MMA warp: {rel(t[1])}")
        for sb in range(min(2 * kb, 32)):
            print(f"      sb {sb:2d}: conv afree-seen {rel(t[80 + 2 * sb])} arrive {rel(t[81 + 2 * sb])} | mma conv-seen {rel(t[2 + 2 * sb])} "
                  f"issued {rel(t[3 + 2 * sb])}")
        print(f"      epilogue of tile 4: tfull {rel(t[160])} cross back {rel(t[161])} main0 loaded {rel(t[164 + 0]) if False else rel(t[160 + 4])} "
              f"[act done {rel(t[170])} staging free {rel(t[171])} stored {rel(t[172])}] main back {rel(t[162])} "
              f"[act done {rel(t[174])} staging free {rel(t[175])} stored {rel(t[176])}] end {rel(t[163])}")
