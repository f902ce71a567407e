#!/usr/bin/env python
"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into one line per kernel launch (the numbers DESIGN.md and
profiles/ quote): duration, tensor-pipe %, DRAM bytes and %, L2 throughput %, L2 hit rate, registers, smem."""
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
units = rows[0]
def g(r, k, d=""):
    return r.get(k, d)
keys = [
    ("dur_us", "gpu__time_duration.sum"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_rt%", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dramR_MB", "dram__bytes_read.sum"),
    ("dramW_MB", "dram__bytes_write.sum"),
    ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2hit%", "lts__t_sector_hit_rate.pct"),
    ("l2_2xbar%", "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("regs", "launch__registers_per_thread"),
    ("warps_act%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("smclk_MHz", "sm__cycles_elapsed.avg.per_second"),
]
print("# unit row:", {k: g(units, c) for k, c in keys})
hdr = f"{'kernel':28s} {'grid':>8s} " + " ".join(f"{k:>10s}" for k, _ in keys)
print(hdr)
for r in rows[1:]:
    name = r["Kernel Name"].split("(")[0].replace("s2s::", "")[:28]
    vals = []
    for k, c in keys:
        v = g(r, c)
        try:
            f = float(v.replace(",", ""))
            u = g(units, c)
            if k.startswith("dram") and k.endswith("MB"):
                f = f / {"byte": 1e6, "Kbyte": 1e3, "Mbyte": 1.0, "Gbyte": 1e-3}.get(u, 1e6)
            if k == "dur_us":
                f = f * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            if k == "smclk_MHz":
                f = f * {"Hz": 1e-6, "Khz": 1e-3, "Mhz": 1.0, "Ghz": 1e3}.get(u, 1e-6)
            vals.append(f"{f:10.1f}")
        except Exception:
            vals.append(f"{v[:10]:>10s}")
    print(f"{name:28s} {r.get('Grid Size', ''):>8s} " + " ".join(vals))
