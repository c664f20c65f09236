"""Summarise an ncu --page raw --csv dump: python tools/ncu_summary.py raw.csv [pattern ...]"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[2:] or [r"gpu__time_duration\.sum", r"dram__bytes_(read|write)\.sum$", r"lts__t_bytes\.sum$", r"l1tex__t_bytes\.sum$",
    r"sm__throughput\.avg\.pct", r"gpu__dram_throughput", r"sm__warps_active\.avg\.pct", r"launch__registers", r"launch__occupancy",
    r"launch__(grid|block)_size", r"smsp__inst_executed\.sum$", r"issue_active\.avg\.pct", r"thread_inst_executed_per_inst", r"sector_hit_rate\.pct$",
    r"smsp__average_warps_issue_stalled.*_per_issue_active", r"sm__inst_executed_pipe_.*\.sum$", r"lts__throughput", r"l1tex__throughput", r"achieved_occupancy",
    r"smsp__cycles_active\.avg$", r"sm__cycles_elapsed\.max", r"local_(load|store)", r"smsp__inst_executed_op_local"]
for i, h in enumerate(hdr):
    if any(re.search(p, h) for p in pats):
        print(f"{h:100s} {units[i]:14s} " + "  ".join(r[i] for r in data))
