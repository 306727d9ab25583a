"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (last pass only)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else len(data) // 2 + 1
agg = collections.OrderedDict()
for r in data[skip:]:
    name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:44]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} n={c:4d} total={t:9.1f}us avg={t / c:8.1f}us share={t / tot * 100:5.1f}%")
print("total us", round(tot, 1), "launches", sum(a[0] for a in agg.values()))
