#!/usr/bin/env python
"""profiles/ncu_traffic.json: mean DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch for each
kernel class the library's profiler reports (bench.py's `roofline.traffic`), over the launches of the LAST steady-state
step of an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log
(tools/ncu_capture.sh).   usage: python tools/ncu_traffic.py gpurun_out/traffic_<tag>.csv > profiles/ncu_traffic.json"""
import csv
import json
import sys

CLASSES = {"tc_gemm_kernel": ("tc_gemm_kernel", "tc_gemm_ts_kernel", "tc_gemm_pair_kernel"), "tc_score_kernel": ("tc_score_kernel", "tc_lse64_kernel"), "tc_emm_pv_kernel": "tc_emm_pv_kernel",
           "la_reduce": ("la_reduce_allheads_kernel", "la_reduce_kv_async_kernel"), "eightpt_kernels": "eightpt",
           "solver": ("ransac_", "pose_select", "essential_to_cand"), "la_apply": "la_apply_allheads_kernel", "la_small_kernel": "la_small",
           "layernorm": "layernorm", "linear_simt_kernel": "linear_simt_kernel", "fpn_fuse": "stem_conv7x7s2",
           "fine_window_gather_kernel": "fine_window_gather_kernel", "fine_match_kernel": "fine_match_kernel"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
path = sys.argv[1]
lines = [ln for ln in open(path) if ln.startswith('"')]
launch = {}
for r in csv.DictReader(lines):
    d = launch.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "bytes": 0.0, "ns": 0.0})
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    if r["Metric Name"].startswith("dram__bytes"):
        d["bytes"] += v
    elif r["Metric Name"] == "gpu__time_duration.sum":
        d["ns"] = v
ids = sorted(launch)
# the last full step: between the last two pos_flatten pairs (two launches per step, at the start of the hand kernels)
marks = [i for i in ids if "pos_flatten" in launch[i]["name"]]
lo, hi = (marks[-4], marks[-2]) if len(marks) >= 4 else (ids[0], ids[-1] + 1)
acc = {}
for i in ids:
    if not (lo <= i < hi):
        continue
    for cls, pat in CLASSES.items():
        pats = pat if isinstance(pat, tuple) else (pat,)
        if any(q in launch[i]["name"] for q in pats):
            a = acc.setdefault(cls, {"bytes": 0.0, "ns": 0.0, "n": 0})
            a["bytes"] += launch[i]["bytes"]
            a["ns"] += launch[i]["ns"]
            a["n"] += 1
out = {k: {"dram_bytes_per_launch": v["bytes"] / v["n"], "launches_in_step": v["n"], "us_per_launch_under_ncu": v["ns"] / v["n"] / 1e3,
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (one steady-state step, all launches), "
                     + path.split("/")[-1]}
       for k, v in acc.items()}
print(json.dumps(out, indent=1))
