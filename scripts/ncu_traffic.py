"""Label an ncu capture of one forward with the library's launch names and write the per-launch DRAM
traffic table bench.py's `roofline.traffic` reads.

    ncu -i prof.ncu-rep --page raw --csv > raw.csv         (here, no GPU needed)
    python scripts/ncu_traffic.py raw.csv marks.txt profiles/r01_traffic_cfg2.json "cfg2 1x3x640x1120"

raw.csv must hold the launches of ONE forward in launch order (ncu -s <launches of the warm-up forwards> -c 89)."""
import csv, json, sys

raw, marks, out, desc = sys.argv[1:5]
rows = list(csv.reader(open(raw)))
h, data = rows[0], rows[2:]
names = [l.split("\t")[0] for l in open(marks).read().strip().splitlines()]
alg = [float(l.split("\t")[2]) for l in open(marks).read().strip().splitlines()]
assert len(names) == len(data), (len(names), len(data))
col = lambda r, k: float(r[h.index(k)])
unit = lambda k: rows[1][h.index(k)]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for nme, a, r in zip(names, alg, data):
    rd = col(r, "dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
    wr = col(r, "dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    e = per.setdefault(nme, {"launches": 0, "dram_bytes": 0.0, "algorithmic_bytes": 0.0, "ncu_time_us": 0.0, "kernel": r[h.index("Kernel Name")].split("(")[0]})
    e["launches"] += 1; e["dram_bytes"] += rd + wr; e["algorithmic_bytes"] += a
    e["ncu_time_us"] += col(r, "gpu__time_duration.sum")
for e in per.values():
    for k in ("dram_bytes", "algorithmic_bytes", "ncu_time_us"):
        e[k] = e[k] / e["launches"]
json.dump({"workload": desc, "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
           "(cold-cache, serialised replays: writes that stay in the 126 MB L2 are not counted)", "per_launch": per},
          open(out, "w"), indent=1)
print("wrote", out, len(per), "kernels")
