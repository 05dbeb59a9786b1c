#!/usr/bin/env python
"""Top source lines / SASS instructions by warp-stall samples for one launch of an .ncu-rep.
usage: python tools/ncu_hot.py rep launch_index [cuda|sass] [topN]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
view = sys.argv[3] if len(sys.argv) > 3 else "cuda"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
if rep.endswith(".csv"):  # exported on the GPU box: ncu -i x.ncu-rep --page source --csv --print-source sass ...
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view, "--launch-skip",
                          str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"#"') or l.startswith('"Line'))
print(lines[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
sc = col["Warp Stall Sampling (All Samples)"]
src = col["Source"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(float(r[sc] or 0) for r in data)
print(f"total samples {tot:.0f}")
data.sort(key=lambda r: -float(r[sc] or 0))
for r in data[:top]:
    s = float(r[sc] or 0)
    reasons = sorted(((float(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    rs = " ".join(f"{n}:{v:.0f}" for v, n in reasons if v > 0)
    ident = r[0]
    print(f"{100 * s / tot:5.1f}%  {ident:>14}  {r[src].strip()[:110]:110s}  [{rs}]")
