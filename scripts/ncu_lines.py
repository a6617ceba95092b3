#!/usr/bin/env python
"""Per-source-line instruction counts of an .ncu-rep in line order, as warp instructions per
`trips` (default: executions of the line with the most... pass the trip count explicitly).
usage: ncu_lines.py report.ncu-rep file_substring first_line last_line [trips]"""
import collections, csv, io, subprocess, sys
rep, fsub, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
trips = float(sys.argv[5]) if len(sys.argv) > 5 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, hd = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
def f(x):
    try: return int(x)
    except Exception: return 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 2 and r[0] == "Line No":
        hd = r; continue
    if hd is None or len(r) < 9 or r[0] == "": continue
    try: ln = int(r[0])
    except Exception: continue
    if fsub not in (cur or "") or ln < l0 or ln > l1: continue
    a = agg[ln]
    a[0] += f(r[hd.index("# Samples")]); a[1] += f(r[hd.index("Instructions Executed")]); a[2] += f(r[hd.index("Thread Instructions Executed")]); a[3] = r[1][:90]
tot = sum(a[1] for a in agg.values())
if trips is None: trips = max(a[1] for a in agg.values())
print(f"lines {l0}-{l1} of {fsub}: {tot} warp instructions = {tot/trips:.1f} per trip ({trips:.3g} trips)")
for ln in sorted(agg):
    a = agg[ln]
    if a[1]: print(f"{ln:5d} {a[1]/trips:6.2f}/trip thr/inst {a[2]/max(a[1],1):5.1f} samp {a[0]:7d} | {a[3]}")
