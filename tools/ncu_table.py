"""One line per kernel from an `ncu --page raw --csv` dump: python tools/ncu_table.py raw.csv
(launches of the same kernel are aggregated: n, total time, and the metrics of its longest launch)."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


units = dict(zip(hdr, rows[1]))
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units.get("gpu__time_duration.sum", "us"), 1.0)


def bytes_of(r, name):
    u = units.get(name, "byte")
    k = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return get(r, name, 0.0) * k


agg = {}
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    t = get(r, "gpu__time_duration.sum", 0.0) * tscale
    a = agg.setdefault(name, {"n": 0, "t": 0.0, "best": None, "bt": -1.0})
    a["n"] += 1
    a["t"] += t
    if t > a["bt"]:
        a["bt"], a["best"] = t, r
print(f"{'kernel (longest launch shown)':58s} {'n':>3s} {'total us':>9s} {'us':>8s} {'DRAM MB':>8s} {'DRAM GB/s':>9s} {'%HBM':>5s} "
      f"{'L2 %pk':>6s} {'L1hit%':>6s} {'L2hit%':>6s} {'issue%':>6s} {'lanes':>5s} {'occ%':>5s} {'regs':>4s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    r = a["best"]
    t = a["bt"]
    dram = bytes_of(r, "dram__bytes_read.sum") + bytes_of(r, "dram__bytes_write.sum")
    if dram == 0.0 and "dram__bytes.sum.per_second" in col:  # section-based captures report the rate only
        u = units.get("dram__bytes.sum.per_second", "byte/s").split("/")[0]
        k = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        dram = get(r, "dram__bytes.sum.per_second", 0.0) * k * t * 1e-6
    print(f"{name[:58]:58s} {a['n']:3d} {a['t']:9.1f} {t:8.1f} {dram / 1e6:8.1f} {dram / max(t, 1e-9) / 1e3:9.1f} "
          f"{get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"{get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{get(r, 'l1tex__t_sector_hit_rate.pct'):6.1f} {get(r, 'lts__t_sector_hit_rate.pct'):6.1f} "
          f"{get(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{get(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} "
          f"{get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} {get(r, 'launch__registers_per_thread'):4.0f}")
