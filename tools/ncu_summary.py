#!/usr/bin/env python
"""Key metrics per profiled launch from an .ncu-rep (read here with `ncu -i ... --page raw --csv`).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront_pct", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("smsp__cycles_active.avg", "smsp_cycles", 1),
    ("sm__cycles_elapsed.max", "sm_cycles", 1),
]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):  # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x_raw.csv`)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    avail = [(k, n, s) for k, n, s in KEYS if k in col]
    print("| # | kernel | grid | block | " + " | ".join(n for _, n, _ in avail) + " |")
    print("|---|---|---|---|" + "---|" * len(avail))
    for r in data:
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("void ", "")[:60]
        vals = []
        for k, n, s in avail:
            v = r[col[k]].replace(",", "")
            try:
                f = float(v)
                u = units[col[k]]
                if n == "time_us":
                    f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                elif n.endswith("_MB"):
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
                vals.append(f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}")
            except ValueError:
                vals.append(v)
        print(f"| {r[col['ID']]} | `{name}` | {r[col['Grid Size']]} | {r[col['Block Size']]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
