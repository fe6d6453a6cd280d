"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
Usage: python tools/launch_agg.py gpurun_out/launches.csv [top]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = r[4].split("(")[0][:60]
    v = float(r[14]) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[13], 1.0)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:62s} n={c:5d} total={t / 1e6:9.3f} ms avg={t / c / 1e3:8.2f} us")
