"""Summarise an `ncu --page source --csv` export: per kernel section, top instructions by stall samples."""
import csv
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in secs:
    hdr = sec["hdr"]
    ix = {h: i for i, h in enumerate(hdr)}
    data = sec["data"]
    f = lambda r, k: int(r[ix[k]] or 0)
    tot = sum(f(r, "# Samples") for r in data)
    toti = sum(f(r, "Instructions Executed") for r in data)
    if tot == 0:
        continue
    print(sec["name"][:90], "samples", tot, "warp-inst", toti)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((sum(f(r, s) for r in data), s) for s in stalls), reverse=True)[:6]
    print("   stall totals:", agg)
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
        st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
        print("  ", r[ix["Address"]][-5:], f(r, "Instructions Executed"), f(r, "# Samples"), r[ix["Source"]][:84], st)
