#!/usr/bin/env python
"""profiles/ncu_traffic.json from exported ncu raw pages: mean DRAM bytes (read + write) per launch for each kernel
class the library's profiler reports (bench.py's `roofline.traffic`).
usage: python tools/ncu_traffic.py gpurun_out/*_raw.csv > profiles/ncu_traffic.json"""
import csv
import io
import json
import sys

CLASSES = {"tc_gemm_kernel": "tc_gemm_kernel", "tc_score_kernel": "tc_score_kernel", "tc_emm_pv_kernel": "tc_emm_pv_kernel",
           "la_reduce": "la_reduce_allheads_kernel", "la_apply": "la_apply_allheads_kernel", "la_small_kernel": "la_small",
           "layernorm": "layernorm", "linear_simt_kernel": "linear_simt_kernel",
           "fine_window_gather_kernel": "fine_window_gather_kernel", "fine_match_kernel": "fine_match_kernel"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(io.StringIO(open(path).read())))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[col["Kernel Name"]]
        for cls, pat in CLASSES.items():
            if pat in name:
                b = 0.0
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    b += float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)
                t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
                a = acc.setdefault(cls, {"bytes": 0.0, "n": 0, "src": set()})
                a["bytes"] += b
                a["n"] += 1
                a["src"].add(path.split("/")[-1])
out = {k: {"dram_bytes_per_launch": v["bytes"] / v["n"], "launches_profiled": v["n"],
           "source": "ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum), " + ", ".join(sorted(v["src"]))}
       for k, v in acc.items()}
print(json.dumps(out, indent=1))
