"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr_i]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = {}
for r in rows[hdr_i + 1:]:
    name = r[ik].split("(")[0].replace("modfx::<unnamed>::", "").replace("void ", "")
    v = float(r[iv].replace(",", ""))
    v *= {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[r[iu]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot * 100:6.1f}%")
