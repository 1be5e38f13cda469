#!/bin/bash
# quick GPU check: full suite + smoke + bench (tag = $1)
set -u
T=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" $O/${T}_pytest.log | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > $O/${T}_bench_cfg2.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python - $T <<'PY'
import json, sys
d = json.loads(open("gpurun_out/%s_bench_cfg2.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],4), d["clocks"]["passes_ms_per_step"], "e2e", round(d["e2e"]["value"],1), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["clocks"]["reasons"])
ks = sorted(d["kernels"], key=lambda k: -k["ms_per_step"])
for k in ks[:40]: print("  %-34s n=%d %6.1f us/launch %6.0f GB/s %5.0f TF" % (k["name"], k["launches_per_step"], 1e3*k["ms_per_step"]/k["launches_per_step"], k["GBps"], k["TFLOPs"]))
PY
