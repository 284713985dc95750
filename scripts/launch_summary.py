"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_summary.py file.csv [last_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            v = v / 1e3 if d["Metric Unit"] == "ns" else v * 1e3 if d["Metric Unit"] == "ms" else v
            agg.setdefault(d["Kernel Name"].split("(")[0][-40:], []).append(v)
tot = 0.0
for k, v in agg.items():
    vv = v[-last:] if last else v
    tot += sum(vv)
    print(f"{k:42s} n={len(vv):4d} sum={sum(vv):10.1f} us   last={vv[-1]:9.1f} us")
print(f"total {tot:.1f} us")
