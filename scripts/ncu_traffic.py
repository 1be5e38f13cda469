"""Label an ncu launch list of one forward with the library's launch names and write (1) the labelled
launch list and (2) the per-launch DRAM traffic table bench.py's `roofline.traffic` reads.

    CIDNET_NO_GRAPH=1 python scripts/prof_forward.py 1 640 1120 3 marks.txt
    CIDNET_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none -s 178 -c 89 --csv --log-file launches.csv python scripts/prof_forward.py 1 640 1120 3
    python scripts/ncu_traffic.py launches.csv marks.txt profiles/r01_traffic_cfg2.json "cfg2 1x3x640x1120" \
        [profiles/r01_launches_cfg2_final.csv]

launches.csv holds 89 consecutive launches of back-to-back identical forwards (ncu's long CSV format: one row
per metric); the window may start anywhere inside a forward, it is rotated so that the stem kernel comes first."""
import csv, json, sys

raw, marks, out, desc = sys.argv[1:5]
labelled = sys.argv[5] if len(sys.argv) > 5 else None
rows = [r for r in csv.reader(l for l in open(raw) if l.startswith('"'))]
h, data = rows[0], rows[1:]
ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Unit", "Metric Value")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
launches = {}
for r in data:
    e = launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]].split("(")[0], "grid": r[ix["Grid Size"]],
                                               "block": r[ix["Block Size"]]})
    e[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
seq = [launches[i] for i in sorted(launches)]
mk = [l.split("\t") for l in open(marks).read().strip().splitlines()]
assert len(seq) == len(mk), (len(seq), len(mk))
first = next(i for i, e in enumerate(seq) if e["kernel"].startswith("stem_"))
seq = seq[first:] + seq[:first]
per = {}
lines = ["launch,name,kernel,grid,block,ncu_time_us,event_time_us,dram_read_bytes,dram_write_bytes,algorithmic_bytes"]
for i, ((name, ev_us, alg), e) in enumerate(zip(mk, seq)):
    rd, wr, t = e["dram__bytes_read.sum"], e["dram__bytes_write.sum"], e["gpu__time_duration.sum"]
    lines.append(f'{i},{name},{e["kernel"]},"{e["grid"]}","{e["block"]}",{t:.2f},{float(ev_us):.1f},{rd:.0f},{wr:.0f},{float(alg):.0f}')
    p = per.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "algorithmic_bytes": 0.0, "ncu_time_us": 0.0, "kernel": e["kernel"]})
    p["launches"] += 1; p["dram_bytes"] += rd + wr; p["algorithmic_bytes"] += float(alg); p["ncu_time_us"] += t
tot = sum(p["ncu_time_us"] for p in per.values())
for p in per.values():
    p["share_of_forward"] = p["ncu_time_us"] / tot
    for k in ("dram_bytes", "algorithmic_bytes", "ncu_time_us"):
        p[k] = p[k] / p["launches"]
json.dump({"workload": desc, "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
           "--clock-control none, one forward launched eagerly (CIDNET_NO_GRAPH=1); values are PER LAUNCH. Serialised "
           "replays: a tensor the previous kernel left in the 126 MB L2 is not re-read from DRAM, and writes that stay in "
           "L2 are not counted, so dram_bytes can be below algorithmic_bytes",
           "ncu_total_us": tot, "per_launch": per}, open(out, "w"), indent=1)
if labelled:
    open(labelled, "w").write("\n".join(lines) + "\n")
print("wrote", out, len(per), "kernels, forward =", round(tot, 1), "us under ncu")
