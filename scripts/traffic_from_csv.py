#!/usr/bin/env python
"""profiles/transport_traffic.json from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` log of `bench.py --steps 1 --warmup 1`: sums over the kernels of the LAST
step (from its largest wf_fly_kernel launch's preceding event kernel to the end).
usage: traffic_from_csv.py log.csv out.json [copy of per-kernel table .csv]"""
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hd = rows[0]
ki, mi, vi, ii = hd.index("Kernel Name"), hd.index("Metric Name"), hd.index("Metric Value"), hd.index("ID")
launches = collections.OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(int(r[ii]), {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
seq = list(launches.values())
unit = {}
for r in rows[1:]:
    unit[r[mi]] = r[hd.index("Metric Unit")]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
fly = [(i, d.get("gpu__time_duration.sum", 0)) for i, d in enumerate(seq) if "wf_fly" in d["name"]]
big = sorted(fly, key=lambda x: -x[1])[:2]
start = max(b[0] for b in big)
while start > 0 and "wf_event_kernel" not in seq[start]["name"]:
    start -= 1
# step back over the ordering kernels of wave 0, if any
while start > 0 and any(k in seq[start - 1]["name"] for k in ("first_nu", "scan_hist", "scatter_order")):
    start -= 1
agg = collections.OrderedDict()
tot = {"r": 0.0, "w": 0.0, "t": 0.0}
for d in seq[start:]:
    if "access_peak" in d["name"]:        # the roofline micro-benchmark of bench.py is not part of the step
        continue
    k = d["name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    r = d.get("dram__bytes_read.sum", 0) * scale.get(unit.get("dram__bytes_read.sum", "byte"), 1.0)
    w = d.get("dram__bytes_write.sum", 0) * scale.get(unit.get("dram__bytes_write.sum", "byte"), 1.0)
    t = d.get("gpu__time_duration.sum", 0) * scale.get(unit.get("gpu__time_duration.sum", "ns"), 1e-6)
    a[0] += 1; a[1] += t; a[2] += r; a[3] += w
    tot["r"] += r; tot["w"] += w; tot["t"] += t
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one bench step "
                 "(1.25e8 packets), all wave-front + fold kernels of the step summed (scripts/traffic_from_csv.py)",
       "dram_bytes_per_launch": tot["r"] + tot["w"], "dram_read": tot["r"], "dram_write": tot["w"], "serialized_ms": tot["t"]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as fh:
        fh.write("kernel,launches,ms,dram_read_GB,dram_write_GB\n")
        for k, a in agg.items():
            fh.write(f"\"{k}\",{a[0]},{a[1]:.3f},{a[2]/1e9:.3f},{a[3]/1e9:.3f}\n")
print(json.dumps(out))
for k, a in agg.items():
    print(f"{a[1]:9.3f} ms {a[2]/1e9:8.2f} + {a[3]/1e9:7.2f} GB  x{a[0]:3d}  {k}")
