"""profiles/traffic.json from the committed ncu summaries (dram__bytes_read.sum + dram__bytes_write.sum per launch): bench.py reads
the file and reports the figure as roofline.traffic with its source, so that no traffic number is ever typed into bench.py.

    python tools/make_traffic_json.py profiles/r02f_wgrad_win_full_summary.csv [more summaries ...]
The FIRST file names the dominant kernel (largest share of a step in the launch list)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        rd = float(r[ix["dram__bytes_read.sum"]]) * UNIT[units[ix["dram__bytes_read.sum"]]]
        wr = float(r[ix["dram__bytes_write.sum"]]) * UNIT[units[ix["dram__bytes_write.sum"]]]
        t = float(r[ix["gpu__time_duration.sum"]]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[units[ix["gpu__time_duration.sum"]]]
        out.append({"kernel": r[ix["Kernel Name"]].split("(")[0], "grid": r[ix["Grid Size"]], "us": t, "dram_read_bytes": rd, "dram_write_bytes": wr})
    return out


files = sys.argv[1:]
per = {os.path.relpath(f, ROOT): load(f) for f in files}
first = per[os.path.relpath(files[0], ROOT)]
mean = sum(e["dram_read_bytes"] + e["dram_write_bytes"] for e in first) / len(first)
doc = {"tc": {"kernel": first[0]["kernel"], "dram_bytes_per_launch": mean, "launches": len(first),
              "source": f"{os.path.relpath(files[0], ROOT)} (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, mean over "
                        f"{len(first)} captured launches of one bench.py run)"},
       "per_launch": per}
json.dump(doc, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(doc["tc"]))
