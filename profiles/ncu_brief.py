#!/usr/bin/env python
"""Brief per-kernel table from an ncu report: python profiles/ncu_brief.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"),
        ("sm__cycles_elapsed.avg.per_second", "GHz")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print("kernel".ljust(44), " ".join(n.rjust(8) for _, n in WANT))
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    vals = []
    for k, n in WANT:
        v = d.get(k, "")
        try:
            f = float(v.replace(",", ""))
            if n == "us":
                f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.get(k), 1.0)
            if n in ("rdMB", "wrMB"):
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u.get(k), 1.0)
            if n == "GHz":
                f *= {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0}.get(u.get(k), 1.0)
            vals.append(f"{f:8.1f}")
        except ValueError:
            vals.append(v[:8].rjust(8))
    name = d.get("Kernel Name", "")
    name = name.replace("void ", "").replace("glf::", "").replace("<unnamed>::", "")
    print(name[:44].ljust(44), " ".join(vals))
