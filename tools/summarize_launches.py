#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (shares per kernel).
usage: python tools/summarize_launches.py launches.csv [skip_first_n] > profiles/xxx.md"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("at::", "").replace("void ", "")
    return name[:110]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        rows.append((int(r["ID"]), short(r["Kernel Name"]), v, r["Grid Size"]))
    rows = [r for r in rows if r[0] >= skip]
    if "--last-step" in sys.argv:  # one steady-state step: between the last two input torch.cat launches
        marks = [i for i, r in enumerate(rows) if "CatArrayBatchedCopy" in r[1] and r[3].startswith("(4736")]
        rows = rows[marks[-2]:marks[-1]]
    for a in sys.argv:   # --last-of=K: the last 1/K of the launches (bench.py runs K equal steps; other workloads)
        if a.startswith("--last-of="):
            k = int(a.split("=")[1])
            rows = rows[len(rows) - len(rows) // k:]
    tot = sum(r[2] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for _, k, v, _g in rows:
        agg[k][0] += 1
        agg[k][1] += v
    print(f"{len(rows)} launches, total device time {tot / 1e6:.2f} ms\n")
    print("| share | total ms | launches | avg us | kernel |")
    print("|---|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot < 0.0005:
            continue
        print(f"| {100 * v / tot:.2f}% | {v / 1e6:.2f} | {n} | {v / n / 1e3:.1f} | `{k}` |")


if __name__ == "__main__":
    main()
