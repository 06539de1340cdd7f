#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python scripts/summarize_ncu.py <tag>        # reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep
"""
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = []

lp = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
lall = os.path.join(ROOT, "gpurun_out", "launches_all_%s.csv" % tag)
if os.path.exists(lall):
    # the whole run (several steps / passes): keep the LAST one, cut at its sample_rays_kernel launch -- the first kernel of the path; the
    # few host-side torch kernels before it (latent mapping) are taken from the previous step so that the list still holds one full step
    rows = [r for r in csv.reader(l for l in open(lall) if not l.startswith("=="))]
    hdr, body = rows[0], [r for r in rows[1:] if len(r) == len(rows[0])]
    ki = hdr.index("Kernel Name")
    marks = [i for i, r in enumerate(body) if "sample_rays_kernel" in r[ki]]
    if len(marks) >= 2:
        per = marks[-1] - marks[-2]
        with open(lp, "w") as fh:
            w = csv.writer(fh, quoting=csv.QUOTE_ALL)
            w.writerow(hdr)
            w.writerows(body[len(body) - per:])
if os.path.exists(lp):
    rows = list(csv.reader(l for l in open(lp) if not l.startswith("==")))
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        name = r[ki].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append("# launch list (%s): per-kernel device time, cold-cache serialised ncu pass -- compare SHARES, not absolutes" % tag)
    out.append("%-92s %7s %12s %7s" % ("kernel", "count", "total_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-92s %7d %12.1f %6.1f%%" % (k, n, t, 100 * t / tot))
    out.append("%-92s %7d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))

rp = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_red.sum",
            "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "smsp__cycles_active.avg", "l1tex__t_bytes.sum", "smsp__inst_executed.sum"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    out.append("")
    out.append("# ncu --set full (%s): selected raw metrics per captured launch (units row: %s)" % (tag, "see ncu raw page"))
    units = rows[1]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out.append("---")
        for w, i in idx:
            out.append("%-70s %s %s" % (w, r[i][:110], units[i]))

os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
dst = os.path.join(ROOT, "profiles", "%s_summary.txt" % tag)
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
