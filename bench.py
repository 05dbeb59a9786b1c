#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the FAR per-pair pose hot path at 640x480 (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference arm: the CPU restatement of the reference
                                                           (oracle/far_oracle.py; the reference is pure Python and
                                                           cannot travel to the GPU box) on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...        one rank per GPU; pairs shard across ranks (weak scaling)

A "step" = one pass of the whole path over one batch of synthetic pairs:
  workload `mp3d_loftr_far` (BASELINE.json configs[1]): FAR-LoFTR forward (ResNet-FPN backbone -> 3x(self,cross)
  linear-attention layers -> dual-softmax coarse matching -> 5x5 fine level) -> weighted 8-point + cheirality
  -> FAR head (regress LoFTR layers -> EMM dual-softmax bilinear attention -> gated pose MLP) -> prior-guided RANSAC
  round (2048 hypotheses / pair, the head's pose as prior) -> FAR head again (gate + blend), batch 32 pairs,
  random-init (seeded) weights, thr = 0 so random features still produce ~1.1k matches/pair.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------- workload
def build_inputs(pairs, seed):
    from far_b200 import synth
    img0, img1 = synth.synth_pair_images(pairs, seed=seed)
    return img0, img1


def cfg_and_weights():
    from far_b200 import synth
    from far_b200.loftr import LoFTR, far_eval_cfg
    cfg = far_eval_cfg(thr=0.0)
    model = LoFTR(cfg)
    sd = synth.synth_state_dict(model.state_dict(), 1234)
    return cfg, model, sd


def flops_emm_call(B, N=4800, h=4, d=64):
    """Algorithmic FLOPs of one far_emm_bilinear_attn call (both directions): S computed twice (LSE pass + recompute
    pass), P V' once, V'^T T once (SURVEY.md 8d 'EMM bilinear attention')."""
    dv = d + 6
    per_dir = 2 * (2.0 * N * N * d) + 2.0 * N * N * dv + 2.0 * N * dv * dv
    return 2 * B * h * per_dir


def cpu_oracle_step(sd, cfg, img0, img1, K):
    """The reference's algorithm for the same workload on the host cores (oracle port), one pair at a time exactly
    like the reference's B=1 evaluation."""
    from oracle import far_oracle as O
    n = img0.shape[0]
    with torch.no_grad():
        for b in range(n):
            d = O.loftr_forward(sd, img0[b:b + 1], img1[b:b + 1], cfg)
            R, t, E = O.pose_from_matches_8pt(d["mkpts0_f"], d["mkpts1_f"], d["mconf"], K, K)
            rt = torch.cat([R, t[:, None]], 1)
            M = int(d["b_ids"].shape[0])
            for _ in range(cfg["fine_pred_steps"]):
                lp, ilp = O.preprocess_helper(rt, M, M, 0, 0)
                O.far_head_mp3d(O._sub(sd, "loftr_regress"), d["featmap0"], d["featmap1"], lp, ilp, cfg)
                R, t, E = O.pose_from_matches_8pt(d["mkpts0_f"], d["mkpts1_f"], d["mconf"], K, K)
    return n


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone runs; each step = a bounded sample (1 pair) of the same workload."""
    if rank != 0:
        return
    from far_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, model, sd = cfg_and_weights()
    del model
    sample_pairs = 1
    img0, img1 = build_inputs(sample_pairs, 20240002)
    K = synth.mp3d_intrinsics(1)[0]
    for _ in range(args.warmup):
        cpu_oracle_step(sd, cfg, img0, img1, K)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(sd, cfg, img0, img1, K)
    dt = time.perf_counter() - t0
    v = sample_pairs * args.steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "image-pairs/sec @640x480", "value": v, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "mp3d_loftr_far", "pairs_per_step": sample_pairs, "thr": 0.0,
                       "note": "CPU restatement of the reference (oracle port), torch CPU fp32, all host threads"},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{sample_pairs} pair(s)/step x {args.steps} steps of the same workload"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_far(args, rank, world, local_rank):
    import torch.distributed as dist
    from far_b200 import synth, ops, _lib
    from far_b200.pipeline import FarPosePipeline, gather_poses
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # the reference's own GPU default: TF32 allowed inside cuDNN convolutions (torch default), matmuls IEEE fp32;
    # every hand-written kernel computes in fp32 (or error-compensated 3xTF32).
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    cfg, model, sd = cfg_and_weights()
    cfg["regress"]["reuse_trunk"] = not args.no_trunk_reuse
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    pairs = args.pairs
    img0_h, img1_h = build_inputs(pairs, 20240002 + rank)
    img0_h, img1_h = img0_h.pin_memory(), img1_h.pin_memory()
    K = synth.mp3d_intrinsics(pairs).to(dev)
    pipe = FarPosePipeline(model, K, K, prior_ransac=not args.no_prior_ransac)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i0, i1):
        return pipe(i0, i1)

    def step_e2e():
        i0 = img0_h.to(dev, non_blocking=True)
        i1 = img1_h.to(dev, non_blocking=True)
        out = pipe(i0, i1)
        return out["pose"].cpu(), out["num_matches"].cpu()  # D2H read of the step's result

    img0_d, img1_d = img0_h.to(dev), img1_h.to(dev)
    for _ in range(args.warmup):
        out = step_resident(img0_d, img1_d)
    barrier()

    # ---------------- timed region 1: inputs resident in HBM (value) ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.timer = ops.OpTimer()
    _lib.profile_enable(True)  # per-kernel CUDA-event pairs inside the library (far_profile_*, include/far_sm100.h)
    launches0 = lib.far_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_resident(img0_d, img1_d)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.far_launch_count() - launches0
    op_times = ops.timer.summary()
    ops.timer = None
    kern = _lib.profile_read()
    _lib.profile_enable(False)
    nmatch = int(out["num_matches"].sum().item())

    # ---------------- timed region 2: end to end through the public API with host buffers ----------------
    step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        pose_h, cnt_h = step_e2e()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    sampler.stop_flag = True

    if world > 1:  # max over ranks + the path's one collective (final pose gather)
        tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
        poses, counts = gather_poses(out["pose"], out["num_matches"])
        assert poses.shape[0] == pairs * world
    if rank != 0:
        return

    hbm, tf, tf_sus, which = load_peaks()
    total_pairs = pairs * world * args.steps
    value = total_pairs / (ms / 1e3)
    e2e = total_pairs / (ms_e2e / 1e3)
    # dominant kernel for the roofline entry: the kernel class with the largest share of the step, timed live by the
    # library's own CUDA events around each of its launches on the launching stream (not the C-ABI call around it)
    top = sorted(op_times.items(), key=lambda kv: -kv[1][1])
    share = {k: round(v[1] / ms, 4) for k, v in top}
    kshare = {k: round(v["ms"] / ms, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
    roof = None
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic_db = json.load(open(tpath))
    if kern:
        name, k = max(kern.items(), key=lambda kv: kv[1]["ms"])
        per_launch_s = k["ms"] / k["launches"] / 1e3
        tensor_bound = name.startswith("tc_")
        if tensor_bound:
            achieved = k["flops"] / k["launches"] / per_launch_s / 1e12
            peak, unit, src = tf_sus, "TFLOP/s", "bf16_tflops_sustained"
        else:
            achieved = k["bytes"] / k["launches"] / per_launch_s / 1e9
            peak, unit, src = hbm, "GB/s", "hbm_gbs"
        tr = traffic_db.get(name)
        roof = {"kernel": name, "bound": "tensor" if tensor_bound else "hbm", "achieved": achieved, "peak": peak,
                "unit": unit, "frac": achieved / peak,
                "peak_source": f"{src} of {which} MEASURED_PEAKS.json (kernel timed inside a long step)",
                "us_per_launch": per_launch_s * 1e6, "launches_per_step": k["launches"] / args.steps,
                "algorithmic_flops_per_launch": k["flops"] / k["launches"],
                "algorithmic_bytes_per_launch": k["bytes"] / k["launches"],
                "traffic": tr["dram_bytes_per_launch"] if tr else None,
                "traffic_source": tr["source"] if tr else None,
                "share_of_step": kshare.get(name),
                "frac_of_3xtf32_ceiling": (achieved / (peak / 6.0)) if tensor_bound else None,
                "note": "3xTF32 error-compensated fp32 GEMM: 3 tensor-core MMAs per algorithmic FMA, so frac <= 1/3 of "
                        "the tf32 pipe (= 1/6 of the bf16 peak used as denominator); traffic = mean DRAM bytes per "
                        "launch over one steady-state step (profiles/ncu_traffic.json)" if tensor_bound else None}
    line = {"metric": "image-pairs/sec @640x480", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "mp3d_loftr_far", "pairs_per_gpu": pairs, "global_pairs": pairs * world,
                       "image": "640x480 gray", "thr": 0.0, "coarse_layers": 3, "fine_pred_steps": 2,
                       "second_solver_call": "prior-guided RANSAC round on the GPU (2048 hypotheses/pair)"
                       if not args.no_prior_ransac else "weighted 8-point + cheirality (as the first call)",
                       "head_trunk": "evaluated once per forward and reused by the 2nd head invocation (identical "
                                     "outputs; --no-trunk-reuse re-evaluates it)" if not args.no_trunk_reuse else
                                     "re-evaluated by each of the 2 head invocations",
                       "matches_per_pair": nmatch / pairs, "parallelism": f"pairs sharded dp{world}",
                       "l2": "working set (inputs 79 MB + weights 204 MB + activations >> 126 MB L2): inputs larger than L2",
                       "backbone": "cuDNN conv with TF32 allowed (the reference's torch default); all other math fp32"},
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * img0_h.numel() * 4 * world),
                    "d2h_bytes_per_step": int((pose_h.numel() * 4 + cnt_h.numel() * 8) * world),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "op_share_of_step": share, "kernel_share_of_step": kshare, "clocks": sampler.summary(), "roofline": roof}
    if world == 1 and not args.no_cpu_baseline:
        cfg2, _, sd2 = cfg_and_weights()
        torch.set_num_threads(os.cpu_count() or 1)
        i0, i1 = build_inputs(1, 20240002)
        Kc = synth.mp3d_intrinsics(1)[0]
        cpu_oracle_step(sd2, cfg2, i0, i1, Kc)  # warm
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 12.0:
            n += cpu_oracle_step(sd2, cfg2, i0, i1, Kc)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n} pair(s) of the same workload, one at a time (reference is batch-1), {dt:.1f} s"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="far", choices=["far", "reference"])
    ap.add_argument("--pairs", type=int, default=32, help="pairs per GPU per step (BASELINE configs[1]: 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prior-ransac", action="store_true",
                    help="re-run the weighted 8-point between the two head invocations instead of the batched GPU "
                         "prior-guided RANSAC round of the FAR recipe (far_b200/ransac.py, 2048 hypotheses per pair)")
    ap.add_argument("--no-trunk-reuse", action="store_true",
                    help="re-evaluate the FAR head trunk in both head invocations, literally as the reference does")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_far(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
