#!/bin/bash
# round-2 evidence run on ONE B200: GPU suite, smoke, bench lines of every single-GPU workload, ncu launch list + DRAM traffic
set -u
O=gpurun_out; T=${1:-r02z}; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" $O/${T}_pytest.log | tail -5
grep -E "hvi margin|LCA[0-9] C=" $O/${T}_pytest.log | head -20
timeout 1800 python -m pytest tests/test_hvi_gpu.py tests/test_hvi_backward_gpu.py tests/test_lca_gpu.py tests/test_zz_bf16_build_gpu.py -m gpu -q -s 2>&1 | grep -E "hvi margin|hvi backward margin|LCA[0-9] C=|bf16 build" > $O/${T}_margins.txt; cat $O/${T}_margins.txt | cut -c1-220
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/${T}_smoke.txt
timeout 600 python bench.py > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference.json 2>/dev/null; echo "bench reference rc=$?"
timeout 600 python bench.py --workload cfg4 --steps 5 > $O/${T}_bench_cfg4.json 2>/dev/null; echo "cfg4 rc=$?"
timeout 600 python bench.py --workload cfg1 > $O/${T}_bench_cfg1.json 2>/dev/null; echo "cfg1 rc=$?"
timeout 600 python bench.py --workload cfg3 > $O/${T}_bench_cfg3.json 2>/dev/null; echo "cfg3 rc=$?"
timeout 600 python bench.py --variant mssa > $O/${T}_bench_cfg2_mssa.json 2>/dev/null; echo "mssa rc=$?"
timeout 600 python bench.py --workload cfg5 --steps 5 > $O/${T}_bench_cfg5_1gpu.json 2>/dev/null; echo "cfg5 1gpu rc=$?"
python - $T <<'PY'
import json, sys
T = sys.argv[1]
for w in ("cfg2", "cfg4", "cfg1", "cfg3", "cfg2_mssa", "cfg5_1gpu", "reference"):
    try:
        d = json.loads(open(f"gpurun_out/{T}_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              "roof", (d["roofline"]["kernel"], round(d["roofline"]["frac"], 3)) if d.get("roofline") else None, d.get("clocks", {}).get("reasons"))
    except Exception as e:
        print(w, "parse failed", e)
PY
export CIDNET_NO_GRAPH=1
timeout 300 python scripts/prof_forward.py 1 640 1120 3 $O/${T}_marks.txt > $O/${T}_prof_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 178 -c 89 --csv \
    --log-file $O/${T}_launches_traffic.csv python scripts/prof_forward.py 1 640 1120 3 > $O/${T}_ncu.log 2>&1
echo "ncu rc=$?"
# per-kernel ncu table of one whole forward of the final tree (reduced section set; summarised ON the box: the report itself is
# too large to travel back)
timeout 900 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section WarpStateStats \
    --section LaunchStats --section Occupancy --section InstructionStats --clock-control none -s 178 -c 89 -o /tmp/${T}_forward -f \
    python scripts/prof_forward.py 1 640 1120 3 > $O/${T}_ncu_sections.log 2>&1
echo "ncu sections rc=$?"
python scripts/ncu_summary.py /tmp/${T}_forward.ncu-rep $O/${T}_marks.txt > $O/${T}_ncu_forward_table.csv 2> $O/${T}_ncu_summary.err; echo "summary rc=$? lines=$(wc -l < $O/${T}_ncu_forward_table.csv)"
rm -f /tmp/${T}_forward.ncu-rep
