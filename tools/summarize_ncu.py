"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares (markdown)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or r[hdr["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[hdr["Kernel Name"]].split("(")[0]
    val = float(r[hdr["Metric Value"]].replace(",", ""))
    unit = r[hdr["Metric Unit"]]
    us = val / 1000.0 if unit in ("nsecond", "ns") else (val if unit in ("usecond", "us") else val * 1000.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
print(f"| **all** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100 % |")
