#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key counters + instruction share per source region.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--lines N]"""
import csv, subprocess, sys, io, re

rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-78s %-12s %s" % (w, units[i], vals[i]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, h2, agg = None, None, []
def I(x):
    try: return int(x)
    except Exception: return 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No": h2 = r; continue
    if h2 and len(r) == len(h2) and r[0] != "":
        d = dict(zip(h2, r))
        agg.append((I(d["Instructions Executed"]), I(d["# Samples"]), cur, int(r[0]), r[1].strip()[:100]))
tot = sum(a[0] for a in agg) or 1
ts = sum(a[1] for a in agg) or 1
print("\ntotal warp instructions %d, stall samples %d" % (tot, ts))
print("top source lines by instructions executed:")
for a in sorted(agg, reverse=True)[:nlines]:
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / ts, a[2], a[3], a[4]))
# regions by function markers in the kernels file
def region_table(fname, path):
    marks = []
    pat = re.compile(r"^\s*(template\s*<[^>]*>\s*)?(__device__|__global__)")
    lines = open(path).read().split("\n")
    name = None
    for n, l in enumerate(lines, 1):
        m = re.search(r"(\w+)\s*\(", l)
        if ("__device__" in l or "__global__" in l) and m:
            # function name = last identifier before '(' on this or the next line
            mm = re.search(r"(\w+)\s*\([^)]*$|(\w+)\s*\(", l.split("__forceinline__")[-1])
            marks.append((n, (mm.group(1) or mm.group(2)) if mm else "?"))
    marks.append((10 ** 9, "end"))
    out = []
    for (a, nm), (b, _) in zip(marks, marks[1:]):
        i = sum(x[0] for x in agg if x[2] == fname and a <= x[3] < b)
        s = sum(x[1] for x in agg if x[2] == fname and a <= x[3] < b)
        if i:
            out.append((100 * i / tot, 100 * s / ts, nm, a))
    return out
import os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print("\ninstruction share per function (source attribution):")
for f in ("bwb_device.cuh", "bwb_kernels.cuh"):
    for pi, ps, nm, a in region_table(f, os.path.join(root, "bwbble_b200", "csrc", f)):
        print("%5.2f%% inst %5.2f%% smp  %s:%s (line %d)" % (pi, ps, f, nm, a))
oth = sum(x[0] for x in agg if x[2] not in ("bwb_device.cuh", "bwb_kernels.cuh"))
print("%5.2f%% inst  other files (intrinsics headers)" % (100 * oth / tot))
