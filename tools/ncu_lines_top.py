"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
share of executed warp instructions and of stall samples.  Usage: ncu_lines_top.py file.csv [kernel indices] [topn]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
secs, cur = [], None


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "lines": []}
        secs.append(cur)
    elif cur is not None and len(r) > 8 and r[0] not in ("Line No", "File Path", ""):
        cur["lines"].append(r)
for k, sec in enumerate(secs):
    if which is not None and k not in which:
        continue
    tot_s = sum(num(r[6]) for r in sec["lines"]) or 1
    tot_i = sum(num(r[7]) for r in sec["lines"]) or 1
    print("====", k, sec["name"][:70], "samples", tot_s, "warp-inst", tot_i)
    for r in sorted(sec["lines"], key=lambda r: -(num(r[7]) / tot_i + num(r[6]) / tot_s))[:topn]:
        print(f"  line {r[0]:>4} inst {num(r[7]) / tot_i * 100:5.1f}% samples {num(r[6]) / tot_s * 100:5.1f}%  {r[1].strip()[:120]}")
