"""Opcode mix from an `ncu --page source --csv --print-source sass` dump: python tools/sass_mix.py src.csv [kernel-index]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# file may hold several kernels: split on "Kernel Name" rows
kernels = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels.append({"name": r[1], "rows": []})
    elif kernels:
        kernels[-1]["rows"].append(r)
k = kernels[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr = k["rows"][0]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix = collections.Counter(); smp = collections.Counter(); total = 0; tot_s = 0
for r in k["rows"][1:]:
    if len(r) <= iex: continue
    src = r[isrc].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    n = int(float(r[iex] or 0)); s = int(float(r[ismp] or 0))
    mix[op] += n; smp[op] += s; total += n; tot_s += s
print(k["name"], "total warp-inst", total, "samples", tot_s)
for op, n in mix.most_common(40):
    print(f"{op:12s} {n:14d} {100*n/total:6.2f}%   samples {100*smp[op]/max(tot_s,1):6.2f}%")
