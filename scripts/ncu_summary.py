#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hottest source lines (needs ncu on PATH).
usage: ncu_summary.py report.ncu-rep [ntop]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_sample_buffer",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_imc_miss",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_drain", "smsp__pcsamp_warps_issue_stalled_mio_throttle"]
print("== raw metrics:", rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued")):
        print(f"{h:86s} {units[i]:16s} {vals[i]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, hd = None, None
agg = collections.defaultdict(lambda: [0, 0, 0])


def f(x):
    try:
        return int(x)
    except Exception:
        return 0


for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hd = r
        continue
    if hd is None or len(r) < 9 or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except Exception:
        continue
    a = agg[(cur, ln, r[1][:80])]
    a[0] += f(r[hd.index("# Samples")]); a[1] += f(r[hd.index("Instructions Executed")]); a[2] += f(r[hd.index("Thread Instructions Executed")])
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"== source: total samples {ts}, warp instructions {ti}")
byfile = collections.defaultdict(lambda: [0, 0])
for (fl, ln, s), a in agg.items():
    byfile[fl][0] += a[0]; byfile[fl][1] += a[1]
for k, v in byfile.items():
    print(f"   {k:32s} samples {100*v[0]/ts:5.1f}%  inst {100*v[1]/ti:5.1f}%")
for (fl, ln, s), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print(f"{fl:14s}{ln:5d} samp {100*a[0]/ts:5.1f}% inst {100*a[1]/ti:5.1f}% thr/inst {a[2]/max(a[1],1):5.1f} | {s}")
