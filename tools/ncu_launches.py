"""Per-launch table from an `ncu --metrics ... --csv --log-file X.csv` run:  python tools/ncu_launches.py X.csv ..."""
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    if not rows:
        print(f, "-- no launches")
        continue
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    by: dict = {}
    for r in rows[1:]:
        by.setdefault((r[ii], r[ki][:48]), {})[r[mi]] = r[vi].replace(",", "")
    print(f)
    for (_, k), m in by.items():
        g = lambda n, d=0.0: float(m.get(n, d) or d)  # noqa: E731
        print(f"  {k:50s} {g('gpu__time_duration.sum') / 1000:8.1f} us  inst {g('smsp__inst_executed.sum') / 1e6:7.1f} M  "
              f"lanes {g('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f}  issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %  "
              f"warps {g('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} %  dram R {g('dram__bytes_read.sum') / 1e6:7.1f} W {g('dram__bytes_write.sum') / 1e6:7.1f} MB  "
              f"regs {m.get('launch__registers_per_thread', '?')}")
