"""Per-kernel summary table of an ncu --set full report: python scripts/ncu_summary.py report.ncu-rep [marks.txt] > table.csv"""
import csv, subprocess, sys
rep = sys.argv[1]
marks = [l.split("\t")[0] for l in open(sys.argv[2]).read().strip().splitlines()] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "us"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor_inst_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefronts_pct"),
        ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__t_bytes.sum", "l2_MB"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg_throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__inst_executed.sum", "warp_inst")]
idx = [(h.index(c), n) for c, n in cols if c in h]
units = rows[1]
w = csv.writer(sys.stdout)
w.writerow((["name"] if marks else []) + [n for _, n in idx])
body = rows[2:]
ki = h.index("Kernel Name")
# align with the marks: the window may start anywhere inside a forward -> find the offset at which kernel types match
def ktype(name):
    for t in ("stem", "head", "dw3x3", "gram", "fold", "iel_gate", "conv_gemm", "sa_", "reduce", "peer"):
        if t in name: return t
    return name
def mtype(m):
    if "stem" in m: return "stem"
    if "head" in m: return "head"
    if "dw3x3" in m: return "dw3x3"
    if "gram" in m: return "gram"
    if "fold" in m: return "fold"
    if "iel_gate" in m: return "iel_gate"
    if m.startswith("sa"): return "sa_"
    return "conv_gemm"
off = 0
if marks:
    best = -1
    for o in range(len(marks)):
        sc = sum(1 for k, r in enumerate(body) if ktype(r[ki]) == mtype(marks[(k + o) % len(marks)]))
        if sc > best: best, off = sc, o
for k, r in enumerate(body):
    vals = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].replace("cidnet::", "").replace("void ", "")[:44]
        else:
            try:
                f = float(v.replace(",", ""))
                if n.endswith("_MB"):
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[i], 1e-6)
                    f *= scale
                if n == "us":
                    f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(units[i], 1.0)
                v = f"{f:.2f}"
            except ValueError:
                pass
        vals.append(v)
    name = [marks[(k + off) % len(marks)]] if marks else []
    w.writerow(name + vals)
