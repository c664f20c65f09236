"""Hot SASS listing from an `ncu --page source --csv --print-source sass` dump: python tools/sass_hot.py src.csv [min_millions]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) * 1e6 if len(sys.argv) > 2 else 10e6
hdr = rows[1]
ia, isrc, iex, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
tot = 0
for r in rows[2:]:
    if len(r) <= iex:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    n = int(float(r[iex] or 0))
    tot += n
    if n >= thr:
        print(f"{a - base:5x} {n / 1e6:8.1f} {int(float(r[ism] or 0)):6d}  {r[isrc].strip()[:100]}")
print("total", tot)
