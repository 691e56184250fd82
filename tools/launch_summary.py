"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path, errors="ignore")))
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    if d["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print(f"total {tot:.1f} us over {n} launches" + (f" = {tot / steps:.1f} us/step, {n / steps:.1f} launches/step" if steps else ""))
print("| launches | total us | us/launch | kernel |\n|---:|---:|---:|---|")
for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | `{name}` |")
