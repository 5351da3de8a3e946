#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list:  python scripts/launch_summary.py launches.csv [frames_per_step steps]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ui, ii = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", "")) * scale[r[ui]]
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["k"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("%-100s %5s %11s %6s %10s %8s" % ("kernel", "n", "total ms", "share", "DRAM GB", "TB/s"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-100s %5d %11.3f %5.1f%% %10.1f %8.2f" % (k[:100], a[0], a[1], 100 * a[1] / tot, a[2], a[2] / a[1] if a[1] else 0))
if len(sys.argv) > 3:
    F, steps = int(sys.argv[2]), int(sys.argv[3])
    ldpc = sum(a[2] for k, a in agg.items() if "ldpc_decode_kernel" in k)
    print("LDPC kernels: %.0f DRAM bytes per frame (%d frames x %d steps)" % (ldpc * 1e9 / F / steps, F, steps))
