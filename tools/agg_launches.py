"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.  usage: agg_launches.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if d["Metric Unit"] in ("us", "usecond"):
        v *= 1e3
    elif d["Metric Unit"] in ("ms", "msecond"):
        v *= 1e6
    agg[d["Kernel Name"][:70]][0] += 1
    agg[d["Kernel Name"][:70]][1] += v
tot = sum(t for _, t in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} n={n:5d} total={t / 1e3:12.1f} us avg={t / n / 1e3:10.1f} us  {100 * t / tot:5.1f}%")
print(f"total {tot / 1e6:.3f} ms")
