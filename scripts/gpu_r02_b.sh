#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
echo "== pytest" ; timeout 1800 python -m pytest tests -m gpu -q --durations=10 > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $O/r02d_pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for pdl in 1 0; do
  echo "== bench PDL=$pdl"; CIDNET_PDL=$pdl timeout 600 python bench.py > $O/r02d_bench_cfg2_pdl$pdl.json 2> $O/r02d_bench_pdl$pdl.err; echo "bench rc=$?"
  python - $pdl <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02d_bench_cfg2_pdl%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], d["clocks"]["passes_ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
    for k in d["kernels"]:
        if "fold" in k["name"] or "gram" in k["name"]: print("  %-34s n=%d %.1f us/launch" % (k["name"], k["launches_per_step"], 1e3*k["ms_per_step"]/k["launches_per_step"]))
except Exception as e:
    print("bench parse failed", e)
PY
done
echo "== cfg4 bench"; timeout 600 python bench.py --workload cfg4 --steps 5 > $O/r02d_bench_cfg4.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02d_bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4', d['value'], d['ms_per_step'], d['clocks'])"
echo "== 3-group experiment"; CIDNET_LIB=$PWD/hvi-cidnet_b200/libcidnet_b200_g3.so timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py tests/test_lca_gpu.py -m gpu -q > $O/r02d_g3_pytest.log 2>&1; echo "g3 rc=$?"; tail -4 $O/r02d_g3_pytest.log
CIDNET_LIB=$PWD/hvi-cidnet_b200/libcidnet_b200_g3.so timeout 300 python bench.py --steps 20 > $O/r02d_bench_g3.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02d_bench_g3.json').read().strip().splitlines()[-1]); print('g3 ms/step', d['ms_per_step'])"
