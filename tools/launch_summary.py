"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py list.csv"""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
agg = {}
for r in rows[1:]:
    t = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:62]:62s} n={n:4d} total={t / 1e3:9.3f} ms  per={t / n:8.1f} us share={100 * t / tot:5.1f}%")
